#!/usr/bin/env python
"""tools/ncu_summary.py -- turn one `ncu --set full --import-source on` capture into the markdown summary kept
under profiles/.  Runs here (no GPU): reads the .ncu-rep with `ncu -i`, disassembles the in-tree library with
nvdisasm to map SASS addresses back to source lines, and attributes executed instructions / stall samples
to the functions of ac_mpc_b200/csrc/mpc_warp.cuh.

  python tools/ncu_summary.py gpurun_out/prof_r1g.ncu-rep > profiles/r1g_ncu_summary.md
"""
import bisect
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ac_mpc_b200", "csrc", "libacmpc_b200.so")
SRC = os.path.join(ROOT, "ac_mpc_b200", "csrc", os.environ.get("NCU_SUMMARY_SRC", "mpc_warp.cuh"))

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
       "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
       "smsp__pcsamp_sample_count"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def disassemble():
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, capture_output=True)
        cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        return subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")


def source_regions():
    src = open(SRC).read().split("\n")
    marks, struct = [], ""
    for i, l in enumerate(src, 1):
        m = re.match(r"struct (\w+)", l)
        if m:
            struct = m.group(1)
        m = re.match(r"(?:template <[^>]*>\s*)?AC_DEV\s+[\w:<>&\s\*]*?\b(\w+)\(", l)
        if m:
            struct = ""
            marks.append((i, m.group(1)))
            continue
        m = re.match(r"\s+AC_MEM\s+(?:explicit\s+)?[\w:<>&\s\*]*?\b(\w+)\(", l)
        if m:
            marks.append((i, (struct + "::" if struct else "") + m.group(1)))
    return marks


def main():
    rep = sys.argv[1]
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    kname = hdr.index("Kernel Name")
    print(f"# ncu --set full summary of `{os.path.basename(rep)}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` on a B200 after the same command "
          "exited 0 without ncu; numbers under ncu are never bench values.\n")
    for vals in raw[2:]:
        print(f"## {vals[kname]}\n\n| metric | value | unit |\n|---|---|---|")
        for h, u, v in zip(hdr, units, vals):
            if h in RAW:
                print(f"| {h} | {v} | {u} |")
        print("\nstall reasons (warp samples, all):\n\n| reason | samples |\n|---|---|")
        st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), int(v.replace(",", "")))
              for h, v in zip(hdr, vals) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued")]
        for h, v in sorted(st, key=lambda t: -t[1]):
            if v:
                print(f"| {h} | {v} |")
        print()
    # per-function attribution through the SASS addresses
    dis = disassemble()
    marks = source_regions()
    starts = [a for a, _ in marks]
    srcpage = ncu_csv(rep, "source")
    # the source page holds one table per kernel
    tables, cur = [], None
    for r in srcpage:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            tables.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for t in tables:
        rows = t["rows"]
        if not rows or t["name"] in seen:
            continue
        seen.add(t["name"])
        h = rows[0]
        ia, ie, isamp, ino = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("stall_no_inst")
        data = [r for r in rows[1:] if len(r) > ino]
        base = int(data[0][ia], 16)
        dyn = {int(r[ia], 16) - base: (int(r[ie]), int(r[isamp]), int(r[ino])) for r in data}
        mang = re.search(r"(acmpc_\w+_kernel)<\(int\)(\d)>", t["name"])
        key = f"{mang.group(1)}ILi{mang.group(2)}E" if mang else None
        infn, chain, fresh, agg = False, [], True, collections.defaultdict(lambda: [0, 0, 0, 0])
        for ln in dis:
            if ln.startswith(".text."):
                infn = key is not None and key in ln
            if not infn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                # `-gi` prints the inline chain innermost frame first, one line per frame, before the instruction
                if fresh:
                    chain, fresh = [], False
                chain.append((os.path.basename(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*)", ln)
            if m:
                fresh = True
                reg = "kernel wrapper (acmpc_b200.cu)"
                for f, l in chain:   # innermost phase-level function of mpc_warp.cuh
                    if f == os.path.basename(SRC):
                        k = bisect.bisect_right(starts, l) - 1
                        name = marks[k][1] if k >= 0 else "?"
                        if "::" in name or name in ("build_waypoints", "speed_instance", "control_instance"):
                            reg = name
                            break
                d = dyn.get(int(m.group(1), 16), (0, 0, 0))
                a = agg[reg]
                a[0] += d[0]; a[1] += d[1]; a[2] += d[2]; a[3] += 1
        tot = [sum(a[i] for a in agg.values()) or 1 for i in range(4)]
        print(f"## where `{t['name'][:70]}` spends its instructions\n")
        print(f"{tot[0]} warp instructions executed, {tot[1]} stall samples, {tot[3]} static SASS instructions. "
              "Attribution: each SASS instruction goes to the innermost SpeedQP / ControlQP method (or phase function) "
              "of its `-lineinfo` inline chain, helpers (shuffles, selects, 3x3 products) included.\n")
        print(f"| function ({os.path.basename(SRC)}) | executed instr. | stall samples | no_instruction samples | static instr. |\n|---|---|---|---|---|")
        for reg, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            if a[0] / tot[0] < 0.004:
                continue
            print(f"| {reg} | {a[0] / tot[0] * 100:.1f} % | {a[1] / tot[1] * 100:.1f} % | {a[2]} | {a[3]} |")
        print()


if __name__ == "__main__":
    main()
