#!/usr/bin/env python
"""Per-opcode and per-source-line view of one kernel of an .ncu-rep captured with
--import-source on: warp instructions and shared-memory wavefronts per unit of work
(e.g. per 1 KiB tile).  Usage: ncu_source_summary.py file.ncu-rep kernel-substring units"""
import csv, subprocess, sys
rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ops, lines, fn, take, hdr, blocks = {}, [], None, False, None, set()
ti = tw = te = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fn = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        take = kern in r[1] and (fn, r[1]) not in blocks      # the page lists every result twice
        blocks.add((fn, r[1]))
        continue
    if r[0] == "Line No":
        hdr = r
        ii, iw, ie = hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
        continue
    if not take or hdr is None:
        continue
    try:
        ie_, iw_, ix_ = int(r[ii]), int(r[iw]), int(r[ie])
    except (ValueError, IndexError):
        continue
    if r[0]:                                    # a CUDA source line: totals of its SASS
        lines.append((fn, int(r[0]), ie_ / units, iw_ / units, ix_ / units, r[1].strip()[:90]))

# opcodes from the plain SASS page (in the combined view a SASS row is listed under every source line it maps to)
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
take, hdr, seen = False, None, 0
for r in csv.reader(sass.splitlines()):
    if not r:
        continue
    if r[0] == "Kernel Name":
        take = kern in r[1]
        seen += take
        take = take and seen == 1            # the page lists every result twice
        continue
    if r[0] == "Address":
        hdr = r
        ii, iw, ie = hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
        continue
    if not take or hdr is None or len(r) < len(hdr):
        continue
    try:
        ie_, iw_, ix_ = int(r[ii]), int(r[iw]), int(r[ie])
    except ValueError:
        continue
    s = r[1].strip()
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    a = ops.setdefault(op, [0, 0, 0])
    a[0] += ie_; a[1] += iw_; a[2] += ix_
    ti += ie_; tw += iw_; te += ix_
print("kernel %s: per unit (%g units): %.0f warp instructions, %.1f shared-memory wavefronts, %.1f of them excessive (bank conflicts)"
      % (kern, units, ti / units, tw / units, te / units))
print("\n%-24s %10s %12s %12s" % ("opcode", "instr/unit", "wavefr/unit", "excess/unit"))
for op, a in sorted(ops.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%-24s %10.1f %12.1f %12.1f" % (op, a[0] / units, a[1] / units, a[2] / units))
print("\nsource lines with >= 4 instructions or >= 2 wavefronts per unit")
for l in lines:
    if l[2] >= 4 or l[3] >= 2:
        print("%-12s %4d  instr %7.1f  wavefr %6.1f  excess %5.1f  %s" % l)
