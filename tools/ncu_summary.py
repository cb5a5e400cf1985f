"""Summarise one kernel launch of an .ncu-rep (captured with `ncu --set full --clock-control none --import-source on`)
into the text + json the judge reads under profiles/:

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_ncu_x.txt [--flops F] [--bytes B] [--shape M,N,K] [--json out.json]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).  States duration, DRAM traffic and achieved
GB/s, tensor-pipe / XU / FMA / issue utilisation, L2 and shared-memory utilisation, registers, occupancy, the stall
reasons of the warp samples, and - with --flops / --bytes - the achieved TFLOP/s or GB/s against the measured peaks."""

import argparse
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("out")
ap.add_argument("--flops", type=float, default=None, help="algorithmic FLOPs of the launch")
ap.add_argument("--bytes", type=float, default=None, help="algorithmic HBM bytes of the launch")
ap.add_argument("--shape", default=None)
ap.add_argument("--json", default=None)
ap.add_argument("--what", default="")
args = ap.parse_args()

raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))


def get(name: str, default=None):
    found = default
    for k, v in d.items():
        if k == name or k.endswith("." + name):
            try:
                return float(v.replace(",", ""))
            except ValueError:
                found = v if v else found
    return found


def to_bytes(name: str) -> float:
    v = get(name, 0.0)
    unit = next((u[k] for k in d if k == name), "byte")
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)


peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
dur_unit = u.get("gpu__time_duration.sum", "us")
dur_us = get("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(dur_unit, 1.0)
dram_r, dram_w = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
lines = [
    f"# {args.what or d.get('Kernel Name', '')}",
    f"kernel            : {d.get('Kernel Name', '')[:150]}",
    f"grid x block      : {d.get('Grid Size', '')} x {d.get('Block Size', '')}, registers/thread {get('launch__registers_per_thread')}, "
    f"dynamic smem {get('launch__shared_mem_per_block_dynamic')} {u.get('launch__shared_mem_per_block_dynamic', '')}",
    f"duration          : {dur_us:.1f} us (under ncu, --clock-control none: cold caches, serialised)",
    f"DRAM traffic      : read {dram_r / 1e6:.2f} MB + write {dram_w / 1e6:.2f} MB = {(dram_r + dram_w) / 1e6:.2f} MB"
    f" -> {(dram_r + dram_w) / dur_us / 1e3:.0f} GB/s = {(dram_r + dram_w) / dur_us / 1e3 / peaks['hbm_gbs']:.2f} of the measured HBM peak ({peaks['hbm_gbs']:.0f} GB/s)",
]
if args.bytes:
    lines.append(f"algorithmic bytes : {args.bytes / 1e6:.2f} MB -> achieved {args.bytes / dur_us / 1e3:.0f} GB/s = {args.bytes / dur_us / 1e3 / peaks['hbm_gbs']:.2f} of the measured HBM peak")
if args.flops:
    tf = args.flops / dur_us / 1e6
    lines.append(f"algorithmic FLOPs : {args.flops / 1e9:.1f} GFLOP -> achieved {tf:.0f} TFLOP/s = {tf / peaks['bf16_tflops']:.2f} of the measured bf16 burst peak ({peaks['bf16_tflops']:.0f} TFLOP/s)")
for label, key in [
    ("tensor pipe active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("XU (MUFU) pipe    ", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("FMA pipe          ", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("ALU pipe          ", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("issue slots active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput     ", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 / smem thr.    ", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 throughput     ", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("DRAM throughput   ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps active      ", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]:
    v = get(key)
    if isinstance(v, float):
        lines.append(f"{label}: {v:.1f} %")
xbar = to_bytes("l1tex__m_xbar2l1tex_read_bytes.sum")
if xbar:
    lines.append(f"L2 -> SM bytes    : {xbar / 1e6:.1f} MB -> {xbar / dur_us / 1e3:.0f} GB/s")
stalls = sorted(((float(v), k.split("issue_stalled_")[1]) for k, v in d.items()
                 if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and v not in ("", "0")), reverse=True)
tot = sum(s for s, _ in stalls) or 1.0
lines.append("warp-state samples: " + ", ".join(f"{name} {100 * s / tot:.0f} %" for s, name in stalls[:8]))
Path(args.out).write_text("\n".join(lines) + "\n")
print("\n".join(lines))
if args.json:
    rec = {"kernel": d.get("Kernel Name", ""), "duration_us": dur_us, "dram_bytes": dram_r + dram_w, "report": Path(args.report).name}
    if args.shape:
        rec["shape"] = [int(x) for x in args.shape.split(",")]
    Path(args.json).write_text(json.dumps(rec) + "\n")
