"""dev tool: attribute ncu per-SASS-instruction counts to CUDA source lines.

  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]

ncu's source page is exported per SASS instruction (offset order); nvdisasm -g on the cubin
of flake_b200/lib/engine.o gives the line of every offset.  Lines are reported twice: by the
innermost (inlined) location and by the outermost caller line inside the kernel.
"""
import csv, io, os, re, subprocess, sys, collections, tempfile
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.environ.get("FB_OBJ", os.path.join(root, "flake_b200", "lib", "engine.o"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], stdout=subprocess.PIPE, text=True).stdout
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
ia, ie, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
isrc = hdr.index("Source")
data = []
for r in rows[2:]:
    if not r or r[0] == 'Kernel Name':
        break
    data.append(r)
base = int(data[0][ia], 16)
# locate the kernel's section in the disassembly by matching the mangled name fragment
short = re.match(r"(?:void )?(\w+)", kname).group(1)
tmpl = re.search(r"<(?:\(int\))?(\d+)>", kname)
sect, cur, name = {}, None, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        name = m.group(1); cur = sect.setdefault(name, []); loc = None; continue
    if cur is None:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inner = (os.path.basename(m.group(1)), int(m.group(2)))
        outer = inner
        for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
            outer = (os.path.basename(mm.group(1)), int(mm.group(2)))
        loc = (inner, outer); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        cur.append((int(m.group(1), 16), loc, m.group(2)))
cands = [k for k in sect if short in k and (tmpl is None or ("ILi%sE" % tmpl.group(1)) in k)]
if not cands:
    raise SystemExit("kernel section not found for " + kname)
sec = {off: (loc, txt) for off, loc, txt in sect[cands[0]]}
inner_cnt, outer_cnt, inner_s, outer_s = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0
for r in data:
    off = int(r[ia], 16) - base
    n = int(r[ie] or 0); s = int(r[isamp] or 0)
    loc = sec.get(off, (None, ""))[0] or (("?", 0), ("?", 0))
    inner_cnt[loc[0]] += n; outer_cnt[loc[1]] += n; inner_s[loc[0]] += s; outer_s[loc[1]] += s
    tot += n; tots += s
print("kernel %s: %d warp instructions, %d stall samples" % (kname[:60], tot, tots))
def src(loc):
    try:
        return open(os.path.join(root, "flake_b200", "csrc", loc[0])).read().splitlines()[loc[1] - 1].strip()[:90]
    except Exception:
        return ""
for title, cnt, smp in (("innermost line", inner_cnt, inner_s), ("outermost line (as seen from the kernel body)", outer_cnt, outer_s)):
    print("---- by %s: inst%% samples%%" % title)
    for loc, n in cnt.most_common(top):
        print("%5.1f%% %5.1f%%  %s:%d  %s" % (100.0 * n / max(1, tot), 100.0 * smp[loc] / max(1, tots), loc[0], loc[1], src(loc)))
if os.environ.get("FB_SASS"):
    print("---- hottest SASS (ncu text | nvdisasm text)")
    hot = sorted(data, key=lambda r: -int(r[ie] or 0))[:int(os.environ["FB_SASS"])]
    for r in hot:
        off = int(r[ia], 16) - base
        loc, txt = sec.get(off, (None, "?"))
        print("%6x %9s  %-50s | %-40s %s" % (off, r[ie], r[isrc].strip()[:50], txt[:40], loc[0] if loc else None))
if os.environ.get("FB_FUNCS"):
    # per device function (noinline callee labels inside the kernel's section): instructions, samples, no_inst stalls
    fn_at, curfn = {}, short
    sec_name = cands[0]
    for ln in dis.splitlines():
        pass
    insec = False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            insec = (m.group(1) == sec_name); curfn = short; continue
        if not insec:
            continue
        m = re.match(r"\s*\.type\s+\$[^$]+\$(\S+),@function", ln)
        if m:
            curfn = m.group(1); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            fn_at[int(m.group(1), 16)] = curfn
    cols = {c: hdr.index(c) for c in ("stall_no_inst", "stall_barrier", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_math", "stall_selected", "stall_branch_resolving")}
    agg = collections.defaultdict(lambda: collections.Counter())
    for r in data:
        off = int(r[ia], 16) - base
        fn = fn_at.get(off, "?")
        a = agg[fn]
        a["inst"] += int(r[ie] or 0); a["samples"] += int(r[isamp] or 0); a["size"] += 1
        for c, i in cols.items():
            a[c] += int(r[i] or 0)
    print("---- per function: size inst%% samples%% | " + " ".join(c.replace("stall_", "") for c in cols))
    for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        print("%5d %5.1f%% %5.1f%% | %s  %s" % (a["size"], 100.0 * a["inst"] / max(1, tot), 100.0 * a["samples"] / max(1, tots),
              " ".join("%6d" % a[c] for c in cols), fn[:70]))
