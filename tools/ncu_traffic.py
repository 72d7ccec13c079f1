"""profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the default kernels on the bench workload, read
from `ncu --set full` reports (no GPU needed here).  bench.py reads the file for `roofline.traffic` / `frac_dram`.

    python tools/ncu_traffic.py profiles/r2_traffic.json cg_stream=gpurun_out/r2j_cg_stream.ncu-rep:4 \
        assembly_gather=gpurun_out/r2j_asm_gather.ncu-rep
`name=report[:iterations]`: every kernel launch in the report is summed; `:k` divides by k PCG iterations per launch."""
import csv
import io
import json
import subprocess
import sys


def dram_bytes(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
            tot += float(r[i].replace(",", "")) * scale
        t = hdr.index("gpu__time_duration.sum")
        tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[t], 1e-6)
        out.append({"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", ""), "dram_bytes": tot,
                    "ms_under_ncu": float(r[t].replace(",", "")) * tscale})
    return out


def main():
    dst, specs = sys.argv[1], sys.argv[2:]
    res = {}
    for spec in specs:
        name, rep = spec.split("=", 1)
        iters = None
        if ":" in rep:
            rep, it = rep.rsplit(":", 1)
            iters = int(it)
        ks = dram_bytes(rep)
        tot = sum(k["dram_bytes"] for k in ks)
        ent = {"source": f"ncu --set full --clock-control none, 1 GPU, cfg 4 ({rep.split('/')[-1]})", "kernels": ks}
        if iters:
            ent["dram_bytes_per_iteration"] = tot / iters
        else:
            ent["dram_bytes_per_assembly"] = tot
        res[name] = ent
    with open(dst, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
