#!/usr/bin/env python3
"""Cut the committed profile summaries out of gpurun_out/ captures.

    python tools/make_profiles.py <tag> <launches.csv> <sweep.ncu-rep> [round-prefix, default r2]

writes profiles/<round>_launches_bench_<tag>_raw.csv, <round>_launches_frame_<tag>.csv (the last steady-state frame of
the launch list: kernel, grid, block, ns, share), <round>_sweep_half_<tag>_ncu_raw.csv / _ncu_summary.txt / .json (the
.json carries the sha of the kernel sources it was measured on: bench.py refuses a stale `traffic`), and
<round>_sweep_half_<tag>_lines.txt (cost per CUDA source line) / _sass_opcodes.txt (opcode histogram of the kernel
and of its inner loops)."""
import csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1:4]
RND = sys.argv[4] if len(sys.argv) > 4 else "r2"
P = os.path.join(ROOT, "profiles")
shutil.copy(launches, os.path.join(P, "%s_launches_bench_%s_raw.csv" % (RND, tag)))
lines = [l for l in open(launches) if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
ker = [(r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"].replace(",", ""))) for r in rows
       if r.get("Metric Name") == "gpu__time_duration.sum"]
# the last complete frame: from the last bond_assign_kernel to the last p2_finish_reduce_kernel
names = [re.sub(r"^void |<.*|\(.*", "", k[0]) for k in ker]
big = max(int(k[1].strip("()").split(",")[0]) for k, n in zip(ker, names) if n == "bond_assign_kernel")
last_end = max(i for i, n in enumerate(names) if n == "p2_finish_reduce_kernel")
last_beg = max(i for i, n in enumerate(names[:last_end]) if n == "bond_assign_kernel")
frame = ker[last_beg:last_end + 1]
tot = sum(k[3] for k in frame)
with open(os.path.join(P, "%s_launches_frame_%s.csv" % (RND, tag)), "w") as f:
    f.write("# ncu launch list of one steady-state frame of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs` (B200)\n")
    f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv ...; per-launch times are cold-cache and serialised\n")
    f.write("kernel,grid,block,gpu__time_duration_ns,share\n")
    for (k, g, b, t), n in zip(frame, names[last_beg:last_end + 1]):
        f.write('"%s",%s,%s,%d,%.3f\n' % (n, g.replace(",", " "), b.replace(",", " "), t, t / tot))
    sw = sum(k[3] for k, n in zip(frame, names[last_beg:]) if n in ("p2_sweep_half_kernel", "p2_fixup_kernel", "p2_sweep_exact_kernel"))
    f.write("# frame total %d ns; K3 (sweep + fixup + overflow check) %d ns = %.3f of the frame\n" % (tot, sw, sw / tot))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, "%s_sweep_half_%s_ncu_raw.csv" % (RND, tag)), "w").write(raw)
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(os.path.join(P, "%s_sweep_half_%s_ncu_summary.txt" % (RND, tag)), "w").write(summ)
r = list(csv.reader(io.StringIO(raw)))
d = dict(zip(r[0], r[2]))
u = dict(zip(r[0], r[1]))
def val(k, scale_unit=True):
    v = float(d[k].replace(",", ""))
    if scale_unit:
        v *= {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(u[k], 1.0)
    return v
t_us = val("gpu__time_duration.sum")
out = {
    "kernel": d["Kernel Name"],
    "source": "profiles/" + RND + "_sweep_half_%s_ncu_raw.csv (ncu --set full --clock-control none, one launch inside bench.py, B200)" % tag,
    "gpu_time_us": t_us,
    "dram_bytes_read": val("dram__bytes_read.sum"),
    "dram_bytes_write": val("dram__bytes_write.sum"),
    "dram_gbs": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / t_us / 1e3,
    "l1tex_throughput_pct": val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    "shared_wavefronts_pct": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    "fma_pipe_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct": val("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warp_instructions": val("smsp__inst_executed.sum"),
    "registers_per_thread": val("launch__registers_per_thread"),
    "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "sm_cycles_active_avg": val("sm__cycles_active.avg", False),
    "sm_cycles_elapsed_max": val("sm__cycles_elapsed.max", False),
}
# the sources the capture belongs to (bench.py: kernel_source_hash())
import hashlib
h = hashlib.sha256()
for name in ("p2_sweep.cu", "half_shell.cuh", "common.cuh"):
    h.update(open(os.path.join(ROOT, "mouse_b200", "csrc", name), "rb").read())
out["kernel_source_sha16"] = h.hexdigest()[:16]
json.dump(out, open(os.path.join(P, "%s_sweep_half.json" % RND), "w"), indent=1)      # the one bench.py reads
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "0.005"], capture_output=True, text=True).stdout
open(os.path.join(P, "%s_sweep_half_%s_lines.txt" % (RND, tag)), "w").write(lines)
regions = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep, "0.006"], capture_output=True, text=True).stdout
open(os.path.join(P, "%s_sweep_half_%s_regions.txt" % (RND, tag)), "w").write(regions)
sass = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_loops.py"), os.path.join(ROOT, "mouse_b200", "libmouse_b200.so"),
                       "p2_sweep_half_kernelILi6ELb0"], capture_output=True, text=True).stdout
sass = "\n".join(l for l in sass.split("\n") if not any(("n=%4d" % k) in l for k in ()) )
open(os.path.join(P, "%s_sweep_half_%s_sass_opcodes.txt" % (RND, tag)), "w").write(
    "# cuobjdump -sass opcode histogram of p2_sweep_half_kernel<6,false> (whole kernel, then every backward-branch loop)\n" + sass)
json.dump(out, open(os.path.join(P, "%s_sweep_half_%s.json" % (RND, tag)), "w"), indent=1)
print(json.dumps(out, indent=1))
print(open(os.path.join(P, "%s_launches_frame_%s.csv" % (RND, tag))).read())
