"""Turn an `ncu --set full` report into the text summary + DRAM-traffic JSON kept under profiles/.
usage: python scripts/profile_summary.py <report.ncu-rep> <summary.txt> <traffic.json> "<workload note>" "<command line>" """
import csv, io, json, subprocess, sys

rep, out_txt, out_json, note, cmdline = sys.argv[1:6]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_max_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
STALLS = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
NAMES = {"adstar_ring_kernel": "Ad_star", "gather3_kernel<0": "Ad_star", "gather3_kernel<(int)0": "Ad_star", "gather3_kernel<1": "compose",
         "gather3_kernel<(int)1": "compose", "compose_ring_kernel": "compose", "slab_fwd": "slab_fwd", "slab_inv": "slab_inv", "xpass": "xpass"}
traffic = {}
with open(out_txt, "w") as f:
    f.write(cmdline + "\n(" + note + "; one launch of each kernel of an EPDiff step)\n\n")
    for r in rows[2:]:
        kn = r[idx["Kernel Name"]]
        f.write("----- %s\n" % kn[:100])
        for w in WANT:
            if w in idx and r[idx[w]] != "":
                f.write("   %-88s %s %s\n" % (w, r[idx[w]], units[idx[w]]))
        top = sorted(((float(r[idx[h]] or 0), h) for h in STALLS), reverse=True)[:5]
        for v, h in top:
            f.write("   %-88s %.3f inst\n" % (h, v))
        key = next((v for k, v in NAMES.items() if k in kn), None)
        if key and key not in traffic:
            def to_bytes(name):
                v, u = float(r[idx[name]]), units[idx[name]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            traffic[key] = {"dram_read_bytes": to_bytes("dram__bytes_read.sum"), "dram_write_bytes": to_bytes("dram__bytes_write.sum"),
                            "workload": note, "source": "ncu --set full --clock-control none, " + out_txt}
json.dump(traffic, open(out_json, "w"), indent=1)
print("wrote", out_txt, out_json, sorted(traffic))
