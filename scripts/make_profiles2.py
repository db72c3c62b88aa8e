#!/usr/bin/env python
"""Round-2 profile summaries under profiles/ from the ncu captures of
    python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline
usage: make_profiles2.py <tag> <launches.csv> <levels.ncu-rep (fast solve: coarsest, level 1, level 0)> [<exact_level0.ncu-rep>]
Writes <tag>_launches.csv, <tag>_lk_levels_ncu_details.txt, <tag>_lk_levels_summary.json (per kernel: duration, DRAM bytes,
L1 data-pipe wavefronts per pixel by source, pipe utilisation, stall reasons per issue), ncu_traffic.json (the level-0 kernel,
what bench.py reports as roofline.traffic) and <tag>_sass_excerpt.txt."""
import csv, io, json, os, re, subprocess, sys

tag, launches, rep = sys.argv[1:4]
rep_exact = sys.argv[4] if len(sys.argv) > 4 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")

rows = [r for r in csv.reader(open(launches, newline="")) if r and not r[0].startswith("==")]
hdr = rows[0]
ik, ig, iv, im = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Name")
with open(os.path.join(out, f"{tag}_launches.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["ID", "Kernel Name", "Grid Size", "gpu__time_duration.sum [ns]"])
    for r in rows[1:]:
        if r[im] == "gpu__time_duration.sum":
            w.writerow([r[0], r[ik], r[ig], r[iv]])


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


det = ncu(rep, "--page", "details")
if rep_exact:
    det += "\n\n==== bit-exact solve, level 0 ====\n" + ncu(rep_exact, "--page", "details")
open(os.path.join(out, f"{tag}_lk_levels_ncu_details.txt"), "w").write(det)


def summaries(rep):
    rr = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    names, units = rr[0], rr[1]
    res = []
    for vals in rr[2:]:
        d, u = dict(zip(names, vals)), dict(zip(names, units))

        def val(k):
            x = float(d[k].replace(",", ""))
            return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "Tbyte": 1e12}.get(u.get(k, ""), 1.0)

        grid = [int(x) for x in re.findall(r"\d+", d["Grid Size"])] if "Grid Size" in d else None
        name = d["Kernel Name"]
        m = re.search(r"lk_level_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(bool\))?(\d), (?:\(bool\))?(\d)(?:, (?:\(bool\))?\d)*>", name)
        dur = val("gpu__time_duration.sum") / {"ns": 1e6, "us": 1e3, "ms": 1.0, "msecond": 1.0, "usecond": 1e3, "nsecond": 1e6}.get(u["gpu__time_duration.sum"], 1e6)
        sh = val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        tot = val("l1tex__data_pipe_lsu_wavefronts.sum") if "l1tex__data_pipe_lsu_wavefronts.sum" in d else None
        res.append({
            "kernel": name.split("(CUtensorMap")[0], "template": m.groups() if m else None, "grid": d.get("Grid Size"),
            "duration_ms_under_ncu": dur,
            "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
            "registers_per_thread": d["launch__registers_per_thread"], "ipc_active": d["sm__inst_executed.avg.per_cycle_active"],
            "warp_inst_executed": d["inst_executed"],
            "l1tex_throughput_pct": d.get("l1tex__throughput.avg.pct_of_peak_sustained_active"),
            "l1_data_pipe_wavefronts": {"shared": sh, "all": tot, "shared_bank_conflicts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")},
            "pipe_pct": {k: d[k] for k in ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                                           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                                           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                                           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                                           "smsp__issue_active.avg.pct_of_peak_sustained_active") if k in d},
            "stalls_per_issue": {k.split("issue_stalled_")[1].split("_per_")[0]: round(float(v), 3) for k, v in d.items()
                                 if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(v) >= 0.05},
        })
    return res


W, H = 1920, 1080
levels = summaries(rep)
PAIRS = int(os.environ.get("PAIRS", "1024"))  # pairs per launch (the bench default; grid.z is larger: the tail pairs of a big batch are 4 CTAs each)
for s in levels:
    if s["template"]:
        mode, cumout = int(s["template"][1]), int(s["template"][2])
        # which pyramid level: coarsest = unwarped; level 1 = warped + cumulative out; level 0 = warped
        gx = int(re.findall(r"\d+", s["grid"])[0])  # column strips: 16 at level 0, 8 at level 1, 4 at level 2 (1080p)
        lvl = {16: 0, 8: 1, 4: 2}[gx]
        px = PAIRS * (W >> lvl) * (H >> lvl)
        targs = re.findall(r"\d+", s["kernel"].split("<")[1].split(">")[0])
        composing = len(targs) >= 6 and targs[5] == "1"
        alg = px * 10 + (0 if mode == 0 else PAIRS * (W >> (lvl + 1)) * (H >> (lvl + 1)) * 8) + (px * 8 if cumout else 0) + \
            (PAIRS * (W >> (lvl + 2)) * (H >> (lvl + 2)) * 8 if composing else 0)
        s.update({"level": lvl, "pixels": px, "algorithmic_bytes_per_launch": alg,
                  "dram_bytes_over_algorithmic": (s["dram_bytes_read"] + s["dram_bytes_write"]) / alg,
                  "thread_inst_per_pixel": float(s["warp_inst_executed"].replace(",", "")) * 32 / px,
                  "shared_wavefronts_per_pixel": s["l1_data_pipe_wavefronts"]["shared"] / px,
                  "l1_wavefronts_per_pixel": (s["l1_data_pipe_wavefronts"]["all"] or 0) / px})
exact = summaries(rep_exact) if rep_exact else []
for s in exact:
    px = PAIRS * W * H
    s.update({"level": 0, "solve": "exact", "pixels": px, "thread_inst_per_pixel": float(s["warp_inst_executed"].replace(",", "")) * 32 / px,
              "shared_wavefronts_per_pixel": s["l1_data_pipe_wavefronts"]["shared"] / px})
json.dump({"fast_solve": levels, "exact_solve_level0": exact}, open(os.path.join(out, f"{tag}_lk_levels_summary.json"), "w"), indent=1)

l0 = [s for s in levels if s.get("level") == 0][0]
j = dict(l0)
j["dram_bytes_per_launch"] = l0["dram_bytes_read"] + l0["dram_bytes_write"]
j["note"] = ("ncu --set full --clock-control none of the level-0 launch (lk_level_kernel<9, bilinear, no cumulative out, fast solve>) of "
             "`python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline` (%d pairs of 1080p per launch). DRAM bytes / algorithmic "
             "bytes = %.3f: no wasted re-reads; the excess is halo rows/columns and the margin of the staged window of next, most of "
             "which L2 absorbs." % (PAIRS, l0["dram_bytes_over_algorithmic"]))
json.dump(j, open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)

# SASS evidence per variant of the 9x9 window
obj = os.path.join(root, "cuda_optical_flow_2_b200", "csrc", "lk_win_9.o")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
pat = re.compile(r"UTMALDG|SYNCS|STG\.E\.ENL2\.256|I2F\.F64|I2F\.S64|MUFU\.RCP64H|MUFU\.RCP\b|IMAD\.WIDE\b|DFMA|DMUL|LDS\.128|BAR\.SYNC")
with open(os.path.join(out, f"{tag}_sass_excerpt.txt"), "w") as f:
    f.write("cuobjdump -sass cuda_optical_flow_2_b200/csrc/lk_win_9.o | grep -E '" + pat.pattern + "' -- counts per kernel variant\n"
            "(template arguments: window, warp mode 0 none / 1 nearest / 2 bilinear, cumulative out, fast solve, fused halo exchange, composing)\n\n")
    cur, counts = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = {}
            continue
        if cur:
            for k in pat.findall(line):
                counts[cur][k] = counts[cur].get(k, 0) + 1
    for fn, c in counts.items():
        m = re.search(r"ILi(\d+)ELi(\d)ELb(\d)ELb(\d)ELb(\d)ELb(\d)E", fn)
        label = f"lk_level_kernel<{m.group(1)},{m.group(2)},{m.group(3)},{m.group(4)},{m.group(5)},{m.group(6)}>" if m else fn
        f.write(f"{label:34s} " + "  ".join(f"{k} {v}" for k, v in sorted(c.items())) + "\n")
    # a few literal lines of the bilinear fast kernel: the TMA loads, the mbarrier wait, a 256-bit store, the conversions
    f.write("\nliteral lines, lk_level_kernel<9,2,0,1,0,1>:\n")
    fn = "_ZN3ofb15lk_level_kernelILi9ELi2ELb0ELb1ELb0ELb1EEEv14CUtensorMap_stS1_S1_S1_NS_14LkKernelParamsE"
    one = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
    seen = set()
    for line in one.splitlines():
        for k in ("UTMALDG", "SYNCS.PHASECHK", "SYNCS.ARRIVE", "STG.E.ENL2.256", "I2F.S64", "MUFU.RCP", "IMAD.WIDE"):
            if k in line and k not in seen:
                seen.add(k)
                f.write(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip() + "\n")
print(json.dumps(j, indent=1)[:1500])
