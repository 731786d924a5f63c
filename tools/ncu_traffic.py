"""Developer tool (no GPU needed): profiles/ncu_traffic.json from an ncu report of the dominant kernel.

    python tools/ncu_traffic.py gpurun_out/prof_ws_r02.ncu-rep "poisson 128^3 nrhs=1" <sweep_bytes stat>

Reads dram__bytes_read.sum + dram__bytes_write.sum per captured launch of wsweep_kernel (ncu --set full,
raw page) and writes the per-launch mean with the kernel name and the packed-factor size the capture
belongs to -- bench.py reports `roofline.traffic` only while both still match the running library."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def main():
    rep, workload, sweep_bytes = sys.argv[1], sys.argv[2], int(sys.argv[3])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ik, ir, iw, it = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                            "gpu__time_duration.sum"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    launches = []
    for r in rows[2:]:
        if "wsweep_kernel" not in r[ik]:
            continue
        b = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
        launches.append({"kernel": r[ik][:80], "dram_bytes": b, "us": float(r[it].replace(",", ""))})
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data[workload] = {"kernel": "wsweep_kernel", "sweep_bytes": sweep_bytes, "launches": launches,
                      "dram_bytes_per_launch": sum(l["dram_bytes"] for l in launches) / max(1, len(launches)),
                      "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, dram__bytes_read.sum + "
                                f"dram__bytes_write.sum, mean of the {len(launches)} captured launches = the sweeps of one apply)"}
    json.dump(data, open(out_path, "w"), indent=1)
    print(json.dumps(data[workload], indent=1)[:600])


if __name__ == "__main__":
    main()
