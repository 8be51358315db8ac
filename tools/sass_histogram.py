"""Per-kernel SASS mnemonic histogram of libvcb200.so (which kernels carry tcgen05 / TMA / TMEM instructions, which run on
mma.sync).  python tools/sass_histogram.py > profiles/r2_sass_histogram.txt   (needs cuobjdump; no GPU)"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

so = Path(__file__).resolve().parents[1] / "video-caption-algorithm_b200" / "csrc" / "libvcb200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "HMMA", "LDGSTS", "UBLKPF", "UBLKCP", "SYNCS", "MUFU",
         "LDG", "STG", "LDS", "STS", "ATOM", "RED", "ACQBULK", "UCGABAR", "BAR"]
kernels: "OrderedDict[str, Counter]" = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name).replace("vc::", "")
        cur = kernels.setdefault(name, Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for w in WATCH:
            if op == w or (w in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "STTM", "HMMA", "MUFU", "UCGABAR", "UBLKPF") and op.startswith(w)):
                cur[w + (".2CTA" if m.group(2) and "2CTA" in m.group(2) else "")] += 1
print(f"{'kernel':64s} total  " + "  ".join(["UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "STTM", "HMMA", "LDGSTS", "UBLKPF", "MUFU", "UCGABAR"]))
for name, c in kernels.items():
    def g(k):
        return c.get(k, 0) + c.get(k + ".2CTA", 0)
    two = "2CTA" if any(k.endswith(".2CTA") for k in c) else ""
    print(f"{name[:64]:64s} {c['total']:5d}  " + "  ".join(f"{g(k):7d}" for k in ["UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "STTM", "HMMA", "LDGSTS", "UBLKPF", "MUFU", "UCGABAR"]) + f"  {two}")
