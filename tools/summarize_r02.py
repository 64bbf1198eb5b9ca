"""Turn the round-2 profiling visit (tools/gpu_round2.sh -> gpurun_out/cap_*.raw.csv, *.hot.txt, launches / traffic csv) into
the tracked summaries under profiles/.  usage: python tools/summarize_r02.py"""
import collections, csv, glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
with open(os.path.join(P, "r02_ncu_full_summary.txt"), "w") as f:
    f.write("# round 2: ncu --set full --clock-control none --import-source on, ONE launch per block, shape-labelled by the command that\n"
            "# produced it (tools/gpu_round2.sh: tools/one_op.py <kind> <args>); values from `ncu -i rep --page raw --csv`, units as ncu prints them.\n"
            "# `us/call` = the same command's CUDA-event time over 20 back-to-back launches WITHOUT ncu (the number to quote).\n")
    for raw in sorted(glob.glob(os.path.join(G, "cap_*.raw.csv"))):
        label = os.path.basename(raw)[4:-8]
        rows = list(csv.reader(open(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        plain = os.path.join(G, "cap_%s.plain.log" % label)
        tail = open(plain).read().strip().splitlines()[-1] if os.path.exists(plain) else ""
        f.write("\n== %s\n   %s\n" % (label, tail))
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            f.write("  kernel: %s\n" % d.get("Kernel Name", "?")[:120])
            for k in KEYS:
                if k in d:
                    f.write("    %-72s %s %s\n" % (k, d[k], u.get(k, "")))
            st = {k: v for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")}
            top = sorted(st.items(), key=lambda kv: -float(kv[1] or 0))[:5]
            f.write("    top stalls: %s\n" % ", ".join("%s=%s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for k, v in top))
        hot = os.path.join(G, "cap_%s.hot.txt" % label)
        if os.path.exists(hot):
            f.write("    hottest SASS instructions (warp-stall samples, tools/ncu_hot.py):\n")
            for line in open(hot).read().splitlines()[:28]:
                f.write("      " + line + "\n")
# launch list
src = os.path.join(G, "launches_unet_b8.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
            k = d["Kernel Name"].split("(")[0][:90]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, "r02_launches_unet_b8.csv"), "w") as f:
        f.write("# one eager SD-1.x UNet call, batch 8, bf16 mode; ncu --metrics gpu__time_duration.sum --clock-control none (tools/gpu_round2.sh)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES\nkernel,launches,total_us,share\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.4f\n' % (k, c, us, us / tot))
        f.write('"TOTAL",%d,%.1f,1.0\n' % (sum(v[0] for v in agg.values()), tot))
# DRAM traffic of the tcgen05 contraction launches of one UNet call
src = os.path.join(G, "traffic_tc_unet_b8.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    hdr, per_id = None, collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                continue
            v = float(d["Metric Value"].replace(",", ""))
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d["Metric Unit"], 1.0)
            per_id[d["ID"]] = per_id.get(d["ID"], 0.0) + v * mult
    if per_id:
        tot = sum(per_id.values())
        json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum of every tc_contract* launch of one eager batch-8 UNet call (ncu, round 2)",
                   "launches": len(per_id), "dram_bytes_total": tot, "dram_bytes_per_launch": tot / len(per_id)},
                  open(os.path.join(P, "r02_tc_traffic.json"), "w"))
print("ok")
