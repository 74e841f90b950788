"""profiles/ncu_traffic.json from `ncu --set full` reports: per kernel (first launch found of each) the DRAM bytes of one
launch and the headline counters, so that bench.py reads `roofline.traffic` from a committed capture of the SAME build
instead of literals.   python profiles/ncu_traffic.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]"""
import csv, io, json, os, subprocess, sys
KEYS = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "launch__registers_per_thread": "registers",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "smsp__inst_executed.sum": "warp_instructions"}
SHORT = ["sphere_trace_kernel", "sdf_forward_ws_kernel", "sdf_forward_tc_kernel", "sdf_backward_mma_kernel", "sdf_backward_tc_kernel",
         "spc_sphere_trace_kernel", "mesh2sdf_dist_kernel", "mesh2sdf_stab_kernel"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3, "ns": 1e-3, "second": 1e6}
out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        short = next((s for s in SHORT if s in name and not (s == "sphere_trace_kernel" and "spc_" in name)), None)
        if short is None:
            continue
        e = {"kernel_name": name[:120], "report": os.path.basename(rep)}
        for k, nice in KEYS.items():
            if k in hdr:
                v = float(row[hdr.index(k)].replace(",", ""))
                u = units[hdr.index(k)]
                if nice.startswith("dram_bytes") or nice == "duration_us":
                    v *= UNIT.get(u, 1)
                e[nice] = v
        out[short] = e
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps({k: (v.get("duration_us"), v.get("dram_bytes_read"), v.get("dram_bytes_write")) for k, v in out.items()}, indent=1))
