"""SASS census of libexl3b200.so: per kernel (template instantiations summed) the instruction count and the tensor-core / TMA mnemonics.
python tools/sass_census.py [lib] > profiles/rNN_sass_census.md"""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "exllamav3_b200/libexl3b200.so"
COLS = ["UTCIMMA", "UTCHMMA", "UTCHMMA.2CTA", "STTM", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UBLKCP", "UTCBAR", "HMMA", "IMMA"]
p = subprocess.Popen(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True)
stats = collections.OrderedDict()
cur = None
for line in p.stdout:
    m = re.search(r"Function : (\S+)", line)
    if m:
        mm = re.match(r"_ZN5exl3b(\d+)", m.group(1))
        name = m.group(1)
        if mm:
            n = int(mm.group(1)); st = m.group(1).index(mm.group(1), 8) + len(mm.group(1)); name = m.group(1)[st:st + n]
        cur = stats.setdefault(name, {"inst": 0, "n": 0, **{c: 0 for c in COLS}})
        cur["n"] += 1
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    cur["inst"] += 1
    base = op.split(".")[0]
    if base in ("UTCHMMA", "UTMALDG") and ".2CTA" in op:
        cur[base + ".2CTA"] += 1
    elif base in cur:
        cur[base] += 1
print(f"# SASS census of `{lib}` (sm_100a only; `cuobjdump -sass`, instantiations of a template summed; tools/sass_census.py)\n")
print("`UTCIMMA` / `UTCHMMA` = `tcgen05.mma` kind::i8 / kind::f16 (`.2CTA` = `cta_group::2`), `STTM` / `LDTM` = `tcgen05.st` / `tcgen05.ld`, "
      "`UTMALDG` = tensor TMA load, `UBLKCP` = bulk copy, `UTCBAR` = `tcgen05.commit`; `HMMA` / `IMMA` (legacy `mma.sync`) do not occur.\n")
print("| kernel | instances | SASS instructions | " + " | ".join(COLS) + " |")
print("|---|---|---|" + "---|" * len(COLS))
tot = {c: 0 for c in COLS}
for k, v in stats.items():
    print(f"| `{k}` | {v['n']} | {v['inst']} | " + " | ".join(str(v[c]) for c in COLS) + " |")
    for c in COLS: tot[c] += v[c]
print("\nTotals: " + ", ".join(f"{c} {tot[c]}" for c in COLS))
