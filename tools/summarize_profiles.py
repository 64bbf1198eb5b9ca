"""Turn the raw artefacts a GPU visit left in gpurun_out/ into the small tracked summaries under profiles/.
usage: python tools/summarize_profiles.py r01"""
import collections, csv, glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

# 1. launch list of one eager UNet call (ncu --metrics gpu__time_duration.sum): aggregate per kernel
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
    with open(os.path.join(P, "%s_launches_unet_b8.csv" % tag), "w") as f:
        f.write("# one eager SD-1.x UNet call, batch 8, bf16 mode; ncu --metrics gpu__time_duration.sum --clock-control none\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES\nkernel,launches,total_us,share\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.4f\n' % (k, c, us, us / tot))
        f.write('"TOTAL",%d,%.1f,1.0\n' % (sum(v[0] for v in agg.values()), tot))

# 2. full captures: key metrics per captured launch
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
with open(os.path.join(P, "%s_ncu_full_summary.txt" % tag), "w") as f:
    f.write("# ncu --set full --clock-control none, one line block per captured launch (raw page); units as ncu prints them\n")
    for rep in sorted(glob.glob(os.path.join(G, "*.ncu-rep"))):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        f.write("\n== %s\n" % os.path.basename(rep))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write("  kernel: %s\n" % d.get("Kernel Name", "?")[:110])
            for k in KEYS:
                if k in d:
                    f.write("    %-72s %s %s\n" % (k, d[k], u.get(k, "")))
            st = {k: v for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")}
            top = sorted(st.items(), key=lambda kv: -float(kv[1] or 0))[:5]
            f.write("    top stalls: %s\n" % ", ".join("%s=%s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for k, v in top))

# 2b. DRAM traffic of every tcgen05 contraction launch of one eager UNet call (ncu dram__bytes_read/write)
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
        json.dump({"kernel": "tc_contract_kernel / tc_contract_pair_kernel", "launches": len(per_id), "dram_bytes_total": tot,
                   "dram_bytes_per_launch": tot / len(per_id),
                   "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one eager SD-1.x UNet call, batch 8 (cold-cache replay per launch)"},
                  open(os.path.join(P, "%s_tc_traffic.json" % tag), "w"))

# 3. bench lines + per-layer table
for name in ("bench_full.log", "bench_ref.log"):
    p = os.path.join(G, name)
    if os.path.exists(p):
        lines = [l for l in open(p) if l.startswith("{")]
        if lines:
            open(os.path.join(P, "%s_%s.json" % (tag, name[:-4])), "w").write(lines[-1])
for nm in ("layers_unet_b8", "layers_vae_b8"):
    p = os.path.join(G, nm + ".log")
    if os.path.exists(p):
        open(os.path.join(P, "%s_%s.txt" % (tag, nm)), "w").write(open(p).read())
print("wrote", sorted(os.listdir(P)))
