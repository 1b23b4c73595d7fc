"""Turns an ncu report into the tracked summaries under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_r1m.ncu-rep profiles/ncu_r1m_kernels.md [profiles/traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

LABELS = {"<2, (int)-1, 2, 1>": "uniform fwd", "<2, 0, 2, 1>": "uniform fwd+bwd STE", "<2, 1, 2, 1>": "uniform fwd+bwd truncated",
          "<2, 2, 2, 1>": "uniform fwd+bwd min/max (headline)", "<3, 4, 2, 1>": "non-uniform fwd K=4 (register table)",
          "points_grad_partial": "centroid gradient K=4", "grid_stats_partial": "grid path (bucket None): chunk min/max",
          "grid_apply": "grid path (bucket None): apply"}
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size"]


def main(rep, out_md, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT]
    mb = lambda v, u: float(v) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]     # noqa: E731
    us = lambda v, u: float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)                     # noqa: E731
    seen, lines = set(), []
    for r in rows[2:]:
        name = r[idx[0]]
        if not any(k in name for k in ("warp_rows_kernel", "points_grad", "grid_")) or name in seen:
            continue
        seen.add(name)
        lines.append([r[i] for i in idx])
    with open(out_md, "w") as f:
        f.write("# ncu --set full --clock-control none (tools/profile_ops.py), 64 Mi float32, s=16, bucket 256 unless noted\n\n"
                "Per launch, first capture of each kernel. Times under ncu are cold-cache and serialised: this table is for DRAM\n"
                "traffic, instruction counts, occupancy and registers; throughput is timed with CUDA events (bench.py, tools/sweep.py).\n\n"
                "| kernel | what | us | DRAM read MB | DRAM write MB | DRAM % of peak | warp instr (M) | warps active % | regs | issue active % | grid |\n"
                "|---|---|---|---|---|---|---|---|---|---|---|\n")
        for v in lines:
            what = next((lab for k, lab in LABELS.items() if k in v[0]), "")
            f.write(f"| `{v[0][:60]}` | {what} | {us(v[1], units[idx[1]]):.1f} | {mb(v[2], units[idx[2]]):.1f} | {mb(v[3], units[idx[3]]):.1f} | "
                    f"{float(v[4]):.1f} | {float(v[5]) / 1e6:.1f} | {float(v[6]):.1f} | {v[7]} | {float(v[8]):.1f} | {v[9]} |\n")
            if traffic_json and "<2, 2, 2, 1>" in v[0]:
                tr = (mb(v[2], units[idx[2]]) + mb(v[3], units[idx[3]])) * 1e6
                json.dump({"uniform_fwd_bwd_minmax_64Mi_dram_bytes": int(tr), "algorithmic_bytes": 16 * (1 << 26),
                           "source": f"{out_md}: dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full)"},
                          open(traffic_json, "w"), indent=1)
    print(open(out_md).read())


if __name__ == "__main__":
    main(*sys.argv[1:])
