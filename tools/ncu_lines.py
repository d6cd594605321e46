#!/usr/bin/env python
"""Per-source-line summary of one kernel from an .ncu-rep (needs -lineinfo and --import-source on):

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_link_planes [top_n]

Prints, per CUDA source line, warp instructions executed, stall samples and excessive shared-memory wavefronts."""
import csv
import io
import subprocess
import sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k",
                          f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    path, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "":
            continue
        d = dict(zip(hdr[4:], r[4:]))

        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        stalls = {k: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        top_stall = max(stalls, key=stalls.get) if stalls else ""
        lines.append((num("Instructions Executed"), num("# Samples"), num("L1 Wavefronts Shared Excessive"),
                      num("L2 Theoretical Sectors Global"), path, r[0], r[1].strip()[:90], top_stall))
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    print(f"total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
    # phases: source lines of the main file grouped by its "// ---- X:" marker comments
    main_file = max(set(l[4] for l in lines), key=lambda f: sum(l[0] for l in lines if l[4] == f))
    try:
        import os
        src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "noether_b200", "csrc",
                                main_file)).read().splitlines()
        marks = [(i + 1, t.strip()[:60]) for i, t in enumerate(src) if t.strip().startswith("// ----")]
        if marks:
            print(f"phases of {main_file}:")
            lo_line = min(int(l[5]) for l in lines if l[4] == main_file)
            bounds = [(lo_line, "(before first marker)")] + marks + [(10**9, "")]
            for (a, name), (b, _) in zip(bounds[:-1], bounds[1:]):
                sel = [l for l in lines if l[4] == main_file and a <= int(l[5]) < b]
                if sel and sum(l[0] for l in sel) > 0:
                    print(f"  {100*sum(l[0] for l in sel)/tot_i:5.1f}% inst {100*sum(l[1] for l in sel)/tot_s:5.1f}% smp  "
                          f"line {a}: {name}")
            oth = [l for l in lines if l[4] != main_file]
            print(f"  {100*sum(l[0] for l in oth)/tot_i:5.1f}% inst {100*sum(l[1] for l in oth)/tot_s:5.1f}% smp  (inlined from other files)")
    except OSError:
        pass
    print("by samples:")
    for l in sorted(lines, key=lambda l: -l[1])[:top]:
        print(f"  {100*l[1]/tot_s:5.1f}% smp {100*l[0]/tot_i:5.1f}% inst  shExc {l[2]:9.0f} L2sec {l[3]:9.0f}  {l[4]}:{l[5]} [{l[7]}] {l[6]}")


if __name__ == "__main__":
    main()
