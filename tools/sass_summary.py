"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md "What proves a Blackwell-native
kernel"): UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), HMMA (legacy mma.sync - expected 0).  Runs here (no GPU): cuobjdump -sass on the built library.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "stc_unet_b200", "libstc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
                                ("UTCBAR", r"\bUTCBAR"), ("UTCATOMSWS(alloc)", r"\bUTCATOMSWS"), ("SYNCS(mbarrier)", r"\bSYNCS"), ("REDG", r"\bRED(G|\.E)"),
                                ("ATOMS", r"\bATOMS"), ("MATCH", r"\bMATCH"), ("LDG.128", r"\bLDG\.E\.128"), ("HMMA(legacy)", r"\bHMMA")])
kern, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        rows[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", line):
        rows[kern]["instr"] += 1
        for name, pat in pats.items():
            if re.search(pat, line):
                rows[kern][name] += 1
dem = subprocess.run(["cu++filt"] + list(rows), capture_output=True, text=True).stdout.splitlines()
names = {k: (d.rsplit("(", 1)[0] if d else k) for k, d in zip(rows, dem)}
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): instruction counts per kernel; kernels listed = those with any async/tensor mnemonic, then totals")
hdr = ["kernel", "instr"] + list(pats)
print(" | ".join(hdr))
tot = collections.Counter()
for k, c in rows.items():
    tot.update(c)
    if any(c[n] for n in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "MATCH")):
        print(" | ".join([names[k][:70], str(c["instr"])] + [str(c[n]) for n in pats]))
print(" | ".join([f"ALL {len(rows)} kernels", str(tot["instr"])] + [str(tot[n]) for n in pats]))
