"""Summarise .ncu-rep captures into a small JSON that is committed under profiles/ and read by bench.py
(`roofline.traffic` must come from a file with provenance, not from a literal).

    python tools/ncu_to_json.py profiles/r2_ncu_metrics.json "<command that was profiled>" gpurun_out/a.ncu-rep [...]
"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    out_path, command, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    try:
        with open(out_path) as f:
            doc = json.load(f)
    except Exception:
        doc = {}
    for path in reps:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1]
            m = {}
            for w in WANT:
                if w in idx:
                    m[w] = {"value": float(r[idx[w]].replace(",", "")) if r[idx[w]] not in ("", "n/a") else None, "unit": units[idx[w]]}
            rd, wr = m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum")
            traffic = None
            if rd and wr and rd["value"] is not None:
                traffic = rd["value"] * UNIT.get(rd["unit"], 1.0) + wr["value"] * UNIT.get(wr["unit"], 1.0)
            doc[name] = {"capture": path.split("/")[-1], "command": command, "dram_traffic_bytes_per_launch": traffic, "metrics": m}
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    print("wrote", out_path, list(doc))


if __name__ == "__main__":
    main()
