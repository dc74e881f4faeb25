"""Per-kernel SASS opcode summary of libyume_b200.so -> profiles/<name>.md (evidence that the hot kernels are
tcgen05 / TMEM / TMA code: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA loads, UTCBAR = tcgen05.commit;
HMMA would be the legacy mma.sync path). Usage: python tools/sass_summary.py [out.md]"""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "yume_b200" / "csrc" / "libyume_b200.so"
WATCH = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "HMMA", "MUFU.EX2", "FMNMX3", "FFMA2", "LDG", "STG", "RED", "SYNCS"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = defaultdict(Counter)
    total = Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    per[cur][w] += 1
    names = demangle(sorted(per.keys() | total.keys()))
    rows = []
    for fn in sorted(total, key=lambda f: -per[f]["UTCHMMA"]):
        short = re.sub(r"\(.*", "", names[fn]).replace("yb::", "").replace("void ", "")
        rows.append((short, total[fn], per[fn]))
    lines = ["# SASS opcode summary of libyume_b200.so (`python tools/sass_summary.py`)", "",
             f"`cuobjdump -sass {LIB.relative_to(ROOT)}`; one row per kernel instance, instruction counts of the mnemonics that "
             "identify the Blackwell paths (B200_PROFILING.md: `tcgen05.mma` -> UTCHMMA, `tcgen05.ld/st` -> LDTM/STTM, TMA -> UTMALDG, "
             "`tcgen05.commit` -> UTCBAR; HMMA would be legacy `mma.sync`).", "",
             "| kernel | instr | " + " | ".join(WATCH) + " |", "|---|---:|" + "---:|" * len(WATCH)]
    for short, tot, c in rows:
        lines.append(f"| `{short}` | {tot} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")
    agg = Counter()
    for _, _, c in rows:
        agg.update(c)
    lines += ["", "Totals: " + ", ".join(f"{w} {agg[w]}" for w in WATCH), ""]
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r02_sass_opcodes.md"
    out.write_text("\n".join(lines))
    print(out, "kernels:", len(rows), "UTCHMMA", agg["UTCHMMA"], "HMMA", agg["HMMA"])


if __name__ == "__main__":
    main()
