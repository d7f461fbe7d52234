"""SASS evidence of what the shipped kernels execute: per kernel of every object in vaevar_b200/build/, the counts of the mnemonics
that identify the Blackwell paths (tcgen05 MMA / TMEM / TMA / packed fp32 / warp-level MMA) -> profiles/<tag>_sass_summary.txt.
    python tools/sass_summary.py r2
Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld,
UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, UTMAPF / UBLKPF = TMA / bulk L2 prefetch, SYNCS = mbarrier, FFMA2 / FMUL2 / FADD2 =
packed fp32, HMMA = mma.sync (fp16 / bf16 / tf32), LDSM = ldmatrix, LDGSTS = cp.async, UCGABAR = cluster barrier."""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKPF", "SYNCS", "UCGABAR", "FFMA2", "FMUL2", "FADD2", "HMMA", "LDSM", "LDGSTS",
        "MUFU", "FFMA", "LDG", "STG", "ATOM", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(tag):
    lines = [__doc__.strip().splitlines()[0], ""]
    for obj in sorted((ROOT / "vaevar_b200" / "build").glob("*.o")):
        sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
        kernels, cur = collections.OrderedDict(), None
        for ln in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                cur = kernels.setdefault(m.group(1), collections.Counter())
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", ln)
            if m and cur is not None:
                op, mods = m.group(1), m.group(2)
                cur["_total"] += 1
                if op in KEYS:
                    cur[op] += 1
                    if op in ("UTCHMMA", "UTMALDG", "UTCBAR") and ".2CTA" in mods:
                        cur[op + ".2CTA"] += 1
        if not kernels:
            continue
        dm = demangle(list(kernels))
        lines.append(f"== {obj.name}: {len(kernels)} kernels")
        for k, c in kernels.items():
            name = re.sub(r"\(.*", "", dm.get(k, k)).replace("void ", "")
            body = " ".join(f"{q}={c[q]}" for q in KEYS + ["UTCHMMA.2CTA", "UTMALDG.2CTA"] if c[q])
            lines.append(f"  {name[:110]:110s} {c['_total']:6d} SASS | {body}")
        lines.append("")
    out = ROOT / "profiles" / f"{tag}_sass_summary.txt"
    out.write_text("\n".join(lines) + "\n")
    print(out, len(lines), "lines")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
