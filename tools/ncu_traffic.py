#!/usr/bin/env python
"""Per-kernel DRAM traffic and time shares from an ncu metrics pass over bench.py:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X.csv python bench.py ...
    python tools/ncu_traffic.py X.csv profiles/ r02
writes <tag>_launch_shares.csv and the <tag>_k_*_traffic.json files bench.py reads for roofline.traffic.
(The absolute times of an ncu pass are cold-cache and serialised; the shares and the DRAM bytes are what is used.)"""
import collections, csv, json, os, re, sys
src, out_dir, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
d = collections.OrderedDict()
for r in rows[1:]:
    rec = dict(zip(hdr, r))
    e = d.setdefault(rec["ID"], {"name": rec["Kernel Name"]})
    e[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
agg = collections.OrderedDict()
for v in d.values():
    n = re.sub(r"\(.*", "", v["name"]).replace("void ", "")
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += v["gpu__time_duration.sum"]; a[2] += v["dram__bytes_read.sum"]; a[3] += v["dram__bytes_write.sum"]
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out_dir, tag + "_launch_shares.csv"), "w") as fh:
    fh.write("kernel,launches,total_ms,share_pct,avg_ms,dram_read_MB,dram_write_MB\n")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        fh.write('"%s",%d,%.3f,%.2f,%.4f,%.1f,%.1f\n' % (n, a[0], a[1] / 1e6, 100 * a[1] / tot, a[1] / a[0] / 1e6, a[2] / 1e6, a[3] / 1e6))


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def launches(pred):
    return [v for v in d.values() if pred(v["name"])]


def traffic(v):
    return v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]


how = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over bench.py (tools/ncu_traffic.py); median over the launches of the kernel"
smem = launches(lambda n: "k_smem<unsigned int, 0>" in n or "k_smem<unsigned int, false>" in n)
small = [v for v in smem if traffic(v) < 0.6 * max(traffic(x) for x in smem)] if smem else []      # the same kernel also runs (plain) on the 1 GB index
if small:
    json.dump({"kernel": "k_smem<u32, plain loads> (100 MB index, L2-resident)", "traffic_bytes_per_launch": med([traffic(v) for v in small]), "reads_per_launch": 2000000,
               "ms_per_launch_under_ncu": med([v["gpu__time_duration.sum"] for v in small]) / 1e6, "launches": len(small), "how": how},
              open(os.path.join(out_dir, tag + "_k_smem_traffic.json"), "w"), indent=1)
pair = launches(lambda n: "k_smem<unsigned int, 1>" in n or "k_smem<unsigned int, true>" in n)
if pair:
    json.dump({"kernel": "k_smem<u32, paired gathers> (1 GB config-3 index, HBM regime)", "traffic_bytes_per_launch": med([traffic(v) for v in pair]), "reads_per_launch": 2000000,
               "ms_per_launch_under_ncu": med([v["gpu__time_duration.sum"] for v in pair]) / 1e6, "launches": len(pair), "how": how},
              open(os.path.join(out_dir, tag + "_k_smem_hbm_traffic.json"), "w"), indent=1)
ov = {k: launches(lambda n, k=k: k in n) for k in ("k_ov_chain", "k_ov_nei", "k_ov_pack")}
if ov["k_ov_chain"]:
    n_seq = int(sys.argv[4]) if len(sys.argv) > 4 else 20000000
    passes = len(ov["k_ov_pack"]) / max(1, round(len(ov["k_ov_pack"]) / max(1, len(ov["k_ov_chain"]))) ) if False else None
    per = {}
    for k, ls in ov.items():
        # the launches of one pass cover n_seq sequences: bytes per sequence = sum over all passes / (passes * n_seq)
        per[k] = sum(traffic(v) for v in ls)
    n_pass = max(1, len(ov["k_ov_pack"]) * 1.0)          # one k_ov_pack per batch; batches per pass = launches / passes
    # batches per pass are not recorded: the caller gives the total number of sequences all captured launches processed
    seqs_total = float(sys.argv[5]) if len(sys.argv) > 5 else None
    if seqs_total:
        json.dump({"kernel": "k_ov_chain<1> + k_ov_nei + k_ov_pack (config-3 index)", "traffic_bytes_per_sequence": sum(per.values()) / seqs_total,
                   "by_kernel_bytes_per_sequence": {k: v / seqs_total for k, v in per.items()}, "sequences_in_capture": seqs_total, "how": how},
                  open(os.path.join(out_dir, tag + "_k_ov_traffic.json"), "w"), indent=1)
print("wrote", [f for f in os.listdir(out_dir) if f.startswith(tag)])
