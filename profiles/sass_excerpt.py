"""SASS evidence for the built library: TMA / bulk-copy / packed-FMA mnemonic counts per kernel and the inner loop of
the dense K=21 filter2d specialisation.   python profiles/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import os
import re
import subprocess
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "trainner_redux_b200", "libotf_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
print(f"# cuobjdump -sass trainner_redux_b200/libotf_b200.so  (sm_100a only: {sorted(set(re.findall(r'arch = (sm_\w+)', txt)))})")
WANT = ("UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "FFMA2", "FFMA", "LDGSTS", "UCGABAR_ARV", "LDS", "STS")
total = Counter()
rows = []
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = dem.replace("void otf::", "").replace("otf::", "").replace("(bool)1", "true").replace("(bool)0", "false")
    dem = re.sub(r">\(.*$", ">", dem) if "<" in dem else re.sub(r"\(.*$", "", dem)
    ops = [re.sub(r"^@!?U?P\d+\s+", "", m.group(1)).split()[0].split(".")[0]
           for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", part)]
    c = Counter(ops)
    total.update(c)
    rows.append((dem, len(ops), c))
print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{w:>8s}" for w in WANT))
for dem, n, c in sorted(rows, key=lambda r: -r[1]):
    if n < 200 and not any(c.get(w) for w in ("UTMALDG", "UBLKCP", "FFMA2")):
        continue
    print(f"{dem[:78]:78s} {n:6d} " + " ".join(f"{c.get(w, 0):8d}" for w in WANT))
print(f"{'TOTAL (all ' + str(len(rows)) + ' kernels)':78s} {sum(r[1] for r in rows):6d} " + " ".join(f"{total.get(w, 0):8d}" for w in WANT))

# the dense K=21 loop of the main filter2d instantiation: the loop body with the most FFMA2
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    if "filter2d_kernelILi8ELi4ELi8ELi16ELb1E" not in part.split("\n", 1)[0]:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", part)]
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            body = [x for x in ins if int(m.group(1), 16) <= x[0] <= a]
            n2 = sum(1 for x in body if "FFMA2" in x[1])
            if best is None or n2 > best[0]:
                best = (n2, body)
    n2, body = best
    c = Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
    print(f"\n# filter2d_kernel<8,4,8,16,true>: dense K=21 row loop, {len(body)} instructions per image row: {dict(c.most_common(8))}")
    print("# first 40 instructions of the loop body:")
    for a, t in body[:40]:
        print(f"    /*{a:05x}*/ {t}")
    # the TMA issue site
    for k, (a, t) in enumerate(ins):
        if "UTMALDG" in t or "UBLKCP" in t:
            print(f"\n# bulk-copy issue site: /*{a:05x}*/ {t}")
