"""Turns an ncu report (or its `--page raw --csv` export) into the tracked summaries under profiles/.

    python tools/summarize_ncu.py <report.ncu-rep | raw.csv> profiles/ncu_<tag>.md [profiles/traffic.json]

Per kernel (first capture of each distinct kernel + grid): time, DRAM bytes read / written, DRAM % of peak,
warp instructions, issue-active %, achieved occupancy, registers, shared memory, and the four largest warp
stall reasons (stalled warps per issue-active cycle).
"""
import csv
import io
import json
import subprocess
import sys

LABELS = {"warp_rows_kernel<2, (int)-1, 2, 1": "uniform fwd, bucket 256", "warp_rows_kernel<2, 0, 2, 1": "uniform fwd+bwd STE, bucket 256",
          "warp_rows_kernel<2, 1, 2, 1": "uniform fwd+bwd truncated, bucket 256",
          "warp_rows_kernel<2, 2, 2, 1, 1>": "uniform fwd+bwd min/max, bucket 256 (HEADLINE; r_b summed per element)",
          "warp_rows_kernel<2, 2, 2, 1, 0>": "uniform bwd min/max alone, bucket 256 (r_b summed in groups)",
          "warp_rows_kernel<2, 2, 2, 1>": "uniform fwd+bwd min/max, bucket 256 (HEADLINE)",
          "warp_rows_kernel<3, 4,": "centroid op K=4 (lane table), bucket 256", "warp_rows_kernel<3, 16,": "centroid op K=16 (lane table), bucket 256",
          "points_grad_partial": "centroid gradient", "grid_stats_partial": "grid path (bucket None): chunk min/max",
          "grid_apply": "grid path (bucket None): apply", "staged_rows_kernel<2, (int)-1": "staged path: uniform fwd",
          "staged_rows_kernel<2, 2": "staged path: fused fwd + min/max bwd", "staged_rows_kernel<3": "staged path: centroid op",
          "plan_sgd_step_kernel": "fused SGD + fix-up + re-quantize (plan)", "plan_nonuniform_fwd_kernel": "centroid plan forward",
          "plan_rows_kernel": "uniform plan", "block_rows_kernel": "block path (round-1 kernel)",
          "unpack_dequant_kernel<1, 4>": "packed codec: 4-bit codes -> dequantized float32 (uniform)",
          "unpack_dequant_kernel<0, 4>": "packed codec: 4-bit codes -> dequantized float32 (centroids)",
          "pack_kernel<4>": "packed codec: uint8 codes -> 4-bit", "inv_scale_kernel": "inverse scaling alone",
          "centroid_index_kernel": "pre-scaled index search"}
BASE = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def load(path):
    if path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    return rows


def main(rep, out_md, traffic_json=None):
    rows = load(rep)
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale_b = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    stall_cols = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]

    def num(r, name):
        try:
            return float(r[col[name]].replace(",", ""))
        except Exception:
            return float("nan")

    seen, out = set(), []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = (name, r[col["launch__grid_size"]] if "launch__grid_size" in col else "")
        if key in seen or not any(k in name for k in ("rows_kernel", "points_grad", "grid_", "plan_", "select_", "pack_", "unpack_", "inv_scale", "centroid_index")):
            continue
        seen.add(key)
        stalls = sorted(((num(r, c), c.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for c in stall_cols), reverse=True)[:4]
        out.append(dict(
            name=name, what=next((lab for k, lab in LABELS.items() if k in name), ""),
            us=num(r, BASE[0]) * scale_t.get(units[col[BASE[0]]], 1.0),
            rd=num(r, BASE[1]) * scale_b.get(units[col[BASE[1]]], 1e-6), wr=num(r, BASE[2]) * scale_b.get(units[col[BASE[2]]], 1e-6),
            dram=num(r, BASE[3]), inst=num(r, BASE[4]) / 1e6, issue=num(r, BASE[5]), occ=num(r, BASE[6]),
            regs=r[col[BASE[7]]], smem=r[col[BASE[8]]] + " " + units[col[BASE[8]]], grid=r[col[BASE[9]]], block=r[col[BASE[10]]],
            stalls=", ".join(f"{n} {v:.2f}" for v, n in stalls)))
    with open(out_md, "w") as f:
        f.write("ncu `--set full --clock-control none`, 64 Mi float32, s=16.  Per launch, first capture of each kernel.  Times under\n"
                "ncu are cold-cache and serialised: this table is for DRAM traffic, instruction counts, occupancy, registers and stall\n"
                "reasons; throughput is timed with CUDA events (bench.py, tools/sweep.py, tools/block_bench.py).\n\n"
                "| kernel | what | us | DRAM rd MB | DRAM wr MB | DRAM % | warp instr M | issue act % | warps act % | regs | dyn smem | grid x block | top stalls (warps per issue-active cycle) |\n"
                "|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for o in out:
            f.write(f"| `{o['name'][:70]}` | {o['what']} | {o['us']:.1f} | {o['rd']:.1f} | {o['wr']:.1f} | {o['dram']:.1f} | {o['inst']:.1f} | "
                    f"{o['issue']:.1f} | {o['occ']:.1f} | {o['regs']} | {o['smem']} | {o['grid']} x {o['block']} | {o['stalls']} |\n")
            if traffic_json and ("warp_rows_kernel<2, 2, 2, 1, 1>" in o["name"] or "warp_rows_kernel<2, 2, 2, 1>" in o["name"]) and o["wr"] > 300:
                json.dump({"uniform_fwd_bwd_minmax_64Mi_dram_bytes": int((o["rd"] + o["wr"]) * 1e6), "algorithmic_bytes": 16 * (1 << 26),
                           "source": f"{out_md} (ncu --set full, one launch of the headline kernel)"}, open(traffic_json, "w"), indent=1)
    print(open(out_md).read())


if __name__ == "__main__":
    main(*sys.argv[1:])
