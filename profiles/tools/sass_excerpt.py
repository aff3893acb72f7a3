"""SASS evidence of the built library (cuobjdump -sass, no GPU needed): per kernel the opcode histogram of the whole
function, and for the kernels named on the command line the full listing.
    python profiles/tools/sass_excerpt.py <liblattigpu.so> <out-prefix> [kernel-substring ...]
Writes <out-prefix>_opcodes.txt (histograms of every ntt_* / ks_* / modup_* kernel, TMA / mbarrier / shuffle mnemonic
counts of the whole library) and <out-prefix>_<name>.sass for each named kernel."""
import collections
import re
import subprocess
import sys

so, prefix = sys.argv[1], sys.argv[2]
want = sys.argv[3:]
raw = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
dem = subprocess.run(["c++filt"], input=raw, capture_output=True, text=True).stdout
funcs = {}
name = None
for line in dem.split("\n"):
    m = re.search(r"Function : (.*)$", line)
    if m:
        name = m.group(1).strip()
        funcs[name] = []
    elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        funcs[name].append(line.rstrip())


def opcode(line):
    body = re.sub(r"/\*.*?\*/", "", line).strip().rstrip(";").strip()
    f = body.split()
    if not f:
        return None
    return f[1] if f[0].startswith("@") and len(f) > 1 else f[0]


tot = collections.Counter()
with open(prefix + "_opcodes.txt", "w") as out:
    out.write("# cuobjdump -sass %s: opcode histograms (static instruction counts)\n" % so)
    for fn, lines in funcs.items():
        ops = collections.Counter(o for o in (opcode(l) for l in lines) if o)
        tot.update(ops)
        if not re.search(r"ntt_|ks_|modup_", fn):
            continue
        out.write("\n%s\n  %d instructions: %s\n" % (fn, sum(ops.values()), ", ".join("%s %d" % kv for kv in ops.most_common(18))))
    keys = [k for k in tot if re.match(r"UTMALDG|UTMASTG|UBLKCP|SYNCS|SHFL|LDGSTS|LDG\.E\.ENL2\.256|STG\.E\.ENL2\.256|CCTL", k)]
    out.write("\n# whole library, memory-movement mnemonics\n")
    for k in sorted(keys):
        out.write("  %-40s %d\n" % (k, tot[k]))
for w in want:
    for fn, lines in funcs.items():
        if w in fn:
            tag = re.sub(r"[^A-Za-z0-9]+", "_", fn.split("::")[-1].split("(")[0]).strip("_")
            with open("%s_%s.sass" % (prefix, tag), "w") as out:
                out.write("// %s\n" % fn)
                out.write("\n".join(lines) + "\n")
