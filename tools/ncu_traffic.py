"""profiles/r2_traffic.json from committed ncu captures: DRAM bytes (read + write) per launch of kernel A.
    python tools/ncu_traffic.py aishell=gpurun_out/r2_prof_fft_aishell.ncu-rep hkust=... libri=..."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for arg in sys.argv[1:]:
    name, rep = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        rd, wr = float(r[col["dram__bytes_read.sum"]]), float(r[col["dram__bytes_write.sum"]])
        unit_r, unit_w = rows[1][col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_write.sum"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        launches.append({"kernel": r[col["Kernel Name"]][:60], "read": rd * scale.get(unit_r, 1), "write": wr * scale.get(unit_w, 1),
                         "us": float(r[col["gpu__time_duration.sum"]]) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(rows[1][col["gpu__time_duration.sum"]], 1)})
    if launches:
        n = len(launches)
        out[name] = {"bytes_per_launch": sum(l["read"] + l["write"] for l in launches) / n,
                     "read_per_launch": sum(l["read"] for l in launches) / n, "write_per_launch": sum(l["write"] for l in launches) / n,
                     "us_per_launch_under_ncu": sum(l["us"] for l in launches) / n, "launches": n, "kernel": launches[0]["kernel"],
                     "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", ".csv") + " (ncu --set full, batches per launch as in bench.py)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
