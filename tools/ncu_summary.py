"""Summarise .ncu-rep files (read here, no GPU needed) into profiles/*.json: per captured kernel the duration,
DRAM bytes, pipe utilisation, L2 -> SM bytes, registers, instruction count and the top stall reasons.
    python tools/ncu_summary.py out.json rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import json
import subprocess
import sys
from collections import Counter

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sectors_srcunit_tex.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                d[h] = f"{v} {u}".strip()
        res.append(d)
    return res


def stalls(rep):
    """Warp-stall sampling per captured kernel (the CSV source page prints every kernel's section twice)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sections, cur, hdr = [], None, None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = (row[1], Counter())
            sections.append(cur)
            hdr = None
        elif row and row[0] == "Address":
            hdr = row
        elif hdr and cur is not None and len(row) == len(hdr):
            for h, v in zip(hdr, row):
                if h.startswith("stall_") and "Not Issued" not in h:
                    try:
                        cur[1][h[6:]] += int(v)
                    except ValueError:
                        pass
    res = []
    for i, (name, c) in enumerate(sections):
        if i > 0 and sections[i - 1][0] == name and sections[i - 1][1] == c:
            continue                                  # the duplicate of the previous section
        tot = sum(c.values()) or 1
        res.append({k: round(100.0 * v / tot, 1) for k, v in c.most_common(6)})
    return res


def main():
    out = {}
    for rep in sys.argv[2:]:
        ks, st = raw(rep), stalls(rep)
        if len(ks) == len(st):
            for k, x in zip(ks, st):
                k["warp_stall_samples_pct_all_warps"] = x
        out[rep.split("/")[-1]] = ks
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:6000])


if __name__ == "__main__":
    main()
