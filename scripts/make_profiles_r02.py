"""Turns the scratch outputs of scripts/gpu_evidence_r02.sh (gpurun_out/r02_*) into the committed summaries under profiles/.

    python scripts/make_profiles_r02.py
"""
import collections
import csv
import gzip
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return None


def dump(name, obj):
    json.dump(obj, open(os.path.join(P, name), "w"), indent=1)
    print("wrote", name)


def ncu_csv_rows(path):
    f = gzip.open(path, "rt", errors="replace") if path.endswith(".gz") else open(path, errors="replace")
    rows = [r for r in csv.reader(f) if len(r) > 5]
    hdr = rows[0]
    return hdr, rows[1:]


def launches_summary():
    src = os.path.join(G, "r02_launches_bench.csv.gz")
    if not os.path.exists(src):
        return
    hdr, rows = ncu_csv_rows(src)
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        v = float(r[i_val].replace(",", ""))
        ms = v / 1e6 if r[i_unit] in ("ns", "nsecond") else v / 1e3 if r[i_unit] in ("us", "usecond") else v
        name = r[i_name].split("(")[0].replace("void ", "").replace("hoig::<unnamed>::", "").replace("unnamed>::", "")[:70]
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + ms)
        total += ms
    with open(os.path.join(P, "r02_launches_bench_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras\n"
                "# (the bench command itself: warm-up + graph capture + 2 timed graph steps + 3 eager roofline steps + 5 e2e steps incl. stage R)\n"
                "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n"
                f"# {len(rows)} launches, {total:.2f} ms total\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{t:10.3f} ms {100 * t / total:5.1f}%  n={n:5d}  {name}\n")
    shutil.copy(src, os.path.join(P, "r02_launches_bench.csv.gz"))
    print("wrote r02_launches_bench_summary.txt")


def conv_traffic():
    src = os.path.join(G, "r02_conv_traffic.csv.gz")
    if not os.path.exists(src):
        return
    hdr, rows = ncu_csv_rows(src)
    i_id, i_metric, i_val, i_unit = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1}
    per = collections.defaultdict(dict)
    for r in rows:
        per[r[i_id]][r[i_metric]] = float(r[i_val].replace(",", "")) * mult.get(r[i_unit], 1)
    n = len(per)
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values())
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values())
    ms = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
    dump("r02_conv_traffic.json", {
        "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'conv_umma_kernel|conv_halo_kernel' "
                  "env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16 (ONE whole forward at batch 64, fp16)",
        "batch": 64, "launches": n, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
        "dram_bytes_per_launch": (rd + wr) / max(n, 1), "conv_ms_under_ncu": ms,
        "note": f"DRAM traffic of all tensor-core conv launches of one forward: {(rd + wr) / 1e9:.1f} GB = {(rd + wr) / 6553e6:.1f} ms at the measured 6.55 TB/s "
                "against ~54 ms of conv time -- the convs are not HBM-bound"})
    shutil.copy(src, os.path.join(P, "r02_conv_traffic.csv.gz"))


COLS = [("gpu__time_duration.sum", "dur"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
        ("dram__bytes.sum.per_second", "DRAM/s"), ("dram__bytes_read.sum", "DRAMrd"), ("dram__bytes_write.sum", "DRAMwr"),
        ("launch__grid_size", "grid"), ("launch__cluster_dim_x", "cl"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_lsb"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar")]


def ncu_table(raw_csvs, out_name, title, top=None):
    """raw_csvs: `ncu -i report --page raw --csv` exports made on the GPU box (the reports themselves are too large to bring back)."""
    rows = []
    for name in raw_csvs:
        path = os.path.join(G, name)
        if not os.path.exists(path):
            continue
        part = list(csv.reader(open(path, errors="replace")))
        h = next(i for i, r in enumerate(part) if r and r[0] == "ID")
        rows = part[h:] if not rows else rows + part[h + 2:]
    if not rows:
        return
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, data = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
    idx = {}
    for key, label in COLS:
        hits = [i for i, n in enumerate(names) if n == key] or [i for i, n in enumerate(names) if n.endswith(key)]
        if hits:
            idx[label] = hits[0]
    i_kn = names.index("Kernel Name")

    def dur_ms(r):
        v, u = float(r[idx["dur"]].replace(",", "")), units[idx["dur"]]
        return v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v if u.startswith("m") else v * 1e3

    order = sorted(range(len(data)), key=lambda i: -dur_ms(data[i]))
    if top:
        order = order[:top]
    with open(os.path.join(P, out_name), "w") as f:
        f.write(f"# {title}\n# ncu --set full --clock-control none (cold cache, serialised replays): one row per kernel instance, sorted by duration"
                + (f" (top {top} of {len(data)})" if top else "") + "\n"
                "# tensor% = sm__pipe_tensor_cycles_active (of active cycles); issue% = issue slots busy; L2% = lts__throughput; DRAM/s = dram__bytes per second (unit in the cell);\n# st_* = average warps stalled per issue-active cycle\n")
        f.write(f"{'kernel':46s} {'ms':>8s} " + " ".join(f"{lab:>9s}" for lab in list(idx)[1:]) + "\n")
        tot = 0.0
        for i in order:
            r = data[i]
            kn = r[i_kn].replace("void ", "").replace("hoig::<unnamed>::", "").split("(hoig")[0].split("(const")[0][:46]
            vals = []
            for lab in list(idx)[1:]:
                v = r[idx[lab]]
                u = units[idx[lab]]
                try:
                    x = float(v.replace(",", ""))
                    vals.append(f"{x:7.1f}{u[:2] if 'byte' in u else '':2s}"[:9].rjust(9) if "/s" not in u else f"{x:5.2f}{u[:2]}/s".rjust(9))
                except ValueError:
                    vals.append(v[:9].rjust(9))
            f.write(f"{kn:46s} {dur_ms(r):8.4f} " + " ".join(vals) + "\n")
            tot += dur_ms(r)
        f.write(f"# listed {tot:.3f} ms\n")
    print("wrote", out_name)


def main():
    os.makedirs(P, exist_ok=True)
    for src, dst in (("r02_bench.log", "r02_bench_n1.json"), ("r02_bench_bf16.log", "r02_bench_n1_bf16.json"),
                     ("r02_bench_ref.log", "r02_bench_reference_arm.json"), ("r02_bench_train.log", "r02_bench_train_n1.json"),
                     ("r02_conditions.log", "r02_condition_stage_b64.json")):
        p = os.path.join(G, src)
        if os.path.exists(p) and last_json(p):
            dump(dst, last_json(p))
    p = os.path.join(G, "r02_prof_convs_b64.log")
    if os.path.exists(p):
        shutil.copy(p, os.path.join(P, "r02_per_conv_shapes_b64.txt"))
    launches_summary()
    conv_traffic()
    ncu_table(["r02_conv_umma_a_raw.csv", "r02_conv_umma_b_raw.csv"], "r02_conv_umma_full.txt",
              "conv_umma_kernel: 88 of the ~125 launches of ONE forward at batch 64 (fp16), all layer families")
    ncu_table(["r02_attn_raw.csv"], "r02_attn_full.txt", "conv_halo_kernel + attn_combine_tc_kernel: the 9 attention layers of one forward at batch 64")
    ncu_table(["r02_ops_raw.csv"], "r02_ops_full.txt", "instnorm_apply / hunfold / hfold / replicate_pad / seg_unfold3: the first 40 of one forward at batch 64")
    ncu_table(["r02_rast_raw.csv"], "r02_rast_full.txt", "rast_bin_kernel + rasterize_kernel, 256 meshes of 13776 faces")


if __name__ == "__main__":
    main()
