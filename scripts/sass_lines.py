"""Joins an ncu SASS page (executed counts) with nvdisasm line info of the same cubin.
    python scripts/sass_lines.py <sass_page.csv> <nvdisasm -g -c output> <mangled kernel substring> [top]
Prints executed warp-instructions per source line (innermost inlined frame) and per opcode."""
import collections, csv, re, sys

page, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows = list(csv.reader(open(page)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
execd = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    execd.append((r[ix['Source']].strip(), int(r[ix['Instructions Executed']] or 0), int(r[ix['# Samples']] or 0)))
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
loc = None; seq = []
for l in lines[start + 1:]:
    if l.startswith("//-----") : break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3)); continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?) ;', l)
    if m: seq.append((m.group(2).strip(), loc))
assert len(seq) <= len(execd) <= len(seq) + 32, (len(seq), len(execd))  # the page also lists the trailing BRA/NOP padding
by_line = collections.Counter(); by_line_op = collections.defaultdict(collections.Counter); samp = collections.Counter()
tot = 0
for (ins, loc), (ins2, n, s) in zip(seq, execd):
    op = [t for t in ins.split() if not t.startswith('@')][0].split('.')[0]
    key = (loc[0], loc[1]) if loc else ("?", 0)
    by_line[key] += n; by_line_op[key][op] += n; samp[key] += s; tot += n
stot = sum(samp.values())
src_cache = {}
def src(f, n):
    import glob
    if f not in src_cache:
        c = glob.glob(f"/root/repo/**/{f}", recursive=True)
        src_cache[f] = open(c[0]).read().split("\n") if c else []
    L = src_cache[f]
    return L[n - 1].strip()[:110] if 0 < n <= len(L) else ""
NORM = float(sys.argv[5]) if len(sys.argv) > 5 else 985600.0
print(f"total {tot}  ({tot/NORM:.1f} per link-warp-substep)")
for key, n in by_line.most_common(top):
    ops = " ".join(f"{o}:{c/NORM:.1f}" for o, c in by_line_op[key].most_common(4))
    print(f"{n/NORM:7.1f} {samp[key]/stot*100:5.1f}%  {key[0]}:{key[1]:<4d} {ops:40s} | {src(*key)}")
