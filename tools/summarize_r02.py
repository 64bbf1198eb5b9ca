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
# bench lines, per-layer tables, library bar, same-box A/B, timelines: copied under their round-2 names
import shutil
for src, dst in (("bench_full.log", "r02_bench_full.json"), ("bench_ref.log", "r02_bench_ref.json"), ("bench_c3.log", "r02_bench_c3.json"),
                 ("bench_c5.log", "r02_bench_c5.json")):
    pth = os.path.join(G, src)
    if os.path.exists(pth):
        lines = [l for l in open(pth).read().splitlines() if l.startswith("{")]
        if lines:
            open(os.path.join(P, dst), "w").write(lines[-1] + "\n")
for src, dst in (("layers_unet_b8.log", "r02_layers_unet_b8.txt"), ("layers_vae_b8.log", "r02_layers_vae_b8.txt"),
                 ("r02_library_bar.txt", "r02_library_bar.txt"), ("ab_unet_step_final.txt", "r02_ab_unet_step.txt"),
                 ("trace_xattn_all.txt", "r02_xattn_timeline.txt"), ("xattn_variants.txt", "r02_xattn_knockouts.txt"),
                 ("attn_ab.txt", "r02_attn_packed_ab.txt")):
    pth = os.path.join(G, src)
    if os.path.exists(pth):
        shutil.copyfile(pth, os.path.join(P, dst))
tl = [os.path.join(G, n) for n in ("trace2_smallk.txt", "trace2_m2048.txt", "trace2_ffout.txt", "trace2_m8192.txt")]
if all(os.path.exists(t) for t in tl):
    with open(os.path.join(P, "r02_pair_timeline.txt"), "w") as f:
        f.write("# round 2: SM-clock timelines of the CTA-pair kernel with the TMA epilogue (tools/trace_pair.py M N K res obf bn; leader CTA of pairs 0 / 36 / 73)\n"
                "# P0 / P1 producer first / last TMA issue, M0w MMA warp reached the unit, M0 accumulator free, M1 first operands landed, M2 last MMA committed\n")
        for t in tl:
            txt = open(t).read()
            f.write(txt[txt.index("args"):] if "args" in txt else txt)
print("copied")
