"""Attribute the warp-stall samples of an `ncu --set full --import-source on` capture to CUDA source lines.

    python tools/ncu_stall_by_line.py <report.ncu-rep> <object.o> <kernel symbol substring> <source file> [top N]

ncu's CLI prints per-SASS-instruction samples but no per-source-line view; nvdisasm -gi gives the (inlined) line of every
SASS offset of the same build.  The two are joined by instruction offset (the report and the object must come from the
same source revision).  Output: samples per source line with the two dominant stall reasons."""
import collections, csv, os, re, subprocess, sys, tempfile


def main():
    rep, obj, sym, src = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
    base_name = os.path.basename(src)
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
        cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and sym in l and l.rstrip().endswith(":"))
    off2line, stack, after_instr = {}, [], True
    for ln in dis[start + 1:]:
        if ln.startswith("//-----") or re.match(r"\s*\.section", ln):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            # a chain of frames, innermost first, each '... inlined at <parent>'; the outermost has no 'inlined at'
            if after_instr:
                stack, after_instr = [], False
            stack.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            after_instr = True
            pick = next((l for f, l in stack if f.endswith(base_name)), None)   # innermost frame inside the requested file
            off2line[int(m.group(1), 16)] = pick
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    names = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    k = next(j for j, i in enumerate(names) if sym.split("ILi")[0].replace("_ZN4cdae", "").strip("0123456789") in rows[i][1] or True)
    h = rows[heads[k]]
    col = {n: i for i, n in enumerate(h)}
    end = names[k + 1] if k + 1 < len(names) else len(rows)
    data = rows[heads[k] + 1:end]
    base = int(data[0][0], 16)
    by, why = collections.Counter(), collections.defaultdict(collections.Counter)
    if os.environ.get("BY_INSTRUCTIONS"):      # executed warp instructions per source line instead of stall samples
        inst = collections.Counter()
        for r in data:
            inst[off2line.get(int(r[0], 16) - base)] += float(r[col["Instructions Executed"]] or 0)
        tot = sum(inst.values())
        text = open(src).read().split("\n")
        print(f"# {rows[names[k]][1]}\n# {tot:.0f} warp instructions executed; source {src}")
        for li, n in inst.most_common(top):
            code = text[li - 1].strip()[:100] if li else "(no line info)"
            print(f"{n:10.0f} {n / tot:6.1%}  L{li}: {code}")
        return
    for r in data:
        s = float(r[col["# Samples"]] or 0)
        if not s:
            continue
        li = off2line.get(int(r[0], 16) - base)
        by[li] += s
        for name, i in col.items():
            if name.startswith("stall_") and "Not Issued" not in name and float(r[i] or 0):
                why[li][name] += float(r[i])
    text = open(src).read().split("\n")
    tot = sum(by.values())
    print(f"# {rows[names[k]][1]}\n# {tot:.0f} samples; source {src}")
    for li, s in by.most_common(top):
        code = text[li - 1].strip()[:100] if li else "(no line info)"
        print(f"{s:6.0f} {s / tot:6.1%}  L{li}: {code}    {dict((a, int(b)) for a, b in why[li].most_common(2))}")


if __name__ == "__main__":
    main()
