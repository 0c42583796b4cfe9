#!/usr/bin/env python3
"""Turn `ncu -i X.ncu-rep --page raw --csv` into a compact per-kernel summary (markdown + json).

    python profiles/summarize_ncu.py gpurun_out/prof_r1.ncu-rep profiles/r1_kernels
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]

SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
         "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def main(rep, out_prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    tensor_cols = [h for h in hdr if "tensor" in h and h.endswith("pct_of_peak_sustained_active") and ".avg." in h]
    out = []
    for r in rows[2:]:
        k = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")}
        for key in KEYS + tensor_cols:
            if key in col and r[col[key]] != "":
                try:
                    v = float(r[col[key]].replace(",", ""))
                except ValueError:
                    continue
                u = units[col[key]]
                if u in SCALE and ("bytes" in key or "duration" in key):
                    v *= SCALE[u]
                    u = "byte" if "bytes" in key else "s"
                if v != 0 or key in KEYS:
                    k[key] = [v, u]
        prev = [o for o in out if o["kernel"] == k["kernel"]]
        if prev:                                  # keep the first launch of a kernel, count the rest
            prev[0]["launches_captured"] = prev[0].get("launches_captured", 1) + 1
            continue
        out.append(k)
    with open(out_prefix + ".json", "w") as f:
        json.dump(out, f, indent=1)
    with open(out_prefix + ".md", "w") as f:
        f.write("# ncu --set full summary of %s\n\n" % rep.split("/")[-1])
        for k in out:
            f.write("## %s\n\n" % k["kernel"])
            for key, v in k.items():
                if key == "launches_captured":
                    f.write("- launches captured: %d (first one shown)\n" % v)
                elif key != "kernel":
                    f.write("- `%s` = %.6g %s\n" % (key, v[0], v[1]))
            if "dram__bytes_read.sum" in k:
                tr = k["dram__bytes_read.sum"][0] + k["dram__bytes_write.sum"][0]
                f.write("- **dram traffic per launch** = %.4g GB; **achieved dram** = %.1f GB/s\n"
                        % (tr / 1e9, tr / k["gpu__time_duration.sum"][0] / 1e9))
            f.write("\n")
    print(open(out_prefix + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
