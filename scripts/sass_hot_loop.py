"""Static instruction count of a kernel's hot loop, read here (no GPU): the longest (or RUN-th longest) branch-free run
of SASS instructions of the function whose mangled name contains `pattern`, as a per-opcode table per `units` units
plus the listing.  The table-mode recursion kernel's branch-free 8-month pass over a lane's two cells is 16 units:
    python scripts/sass_hot_loop.py gr2msemidistr_b200/libgr2m_b200.so _Z13k_gr2m_calib4ILb0ELi3ELi2ELi3EE 16 > profiles/x.txt
The dispatch model of DESIGN 4/K1 (2 issue cycles per FP64-pipe instruction, 1 per other) is printed with it."""
import collections, os, re, subprocess, sys

so, pat = sys.argv[1], sys.argv[2]
units = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ins, on = [], False
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
    elif on:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
op = lambda t: (t.split()[1] if t.startswith("@") else t.split()[0])
runs, s0 = [], None
for i, (_, t) in enumerate(ins):
    if op(t).split(".")[0] in ("BRA", "EXIT", "RET", "CALL", "BRX", "JMP"):
        if s0 is not None:
            runs.append((s0, i))
        s0 = i + 1
    elif s0 is None:
        s0 = i
runs.sort(key=lambda r: r[1] - r[0], reverse=True)
a, b = runs[int(os.environ.get("RUN", "0"))]
cnt = collections.Counter(op(t) for _, t in ins[a:b])
n = b - a
fp64 = sum(v for k, v in cnt.items() if k.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"# {pat}: {len(ins)} instructions; branch-free run {ins[a][0]:#x}..{ins[b - 1][0]:#x}: {n} instructions = "
      f"{n / units:.2f} per unit, FP64 pipe {fp64} = {fp64 / units:.2f} per unit, "
      f"issue cycles per unit (2 per FP64, 1 per other) = {(n + fp64) / units:.2f}")
print(f"# {'opcode':28s}{'per ' + str(units) + ' units':>14s}{'per unit':>10s}")
for k, v in cnt.most_common():
    print(f"# {k:28s}{v:14d}{v / units:10.2f}")
if not os.environ.get("NO_LISTING"):
    print("# ---- listing ----")
    for adr, t in ins[a:b]:
        print(f"/*{adr:04x}*/  {t}")
