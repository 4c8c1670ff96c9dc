#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into the text files kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/summary.txt

Per kernel launch in the report: duration, DRAM bytes, L2 / shared-memory / issue utilisation,
the warp-stall breakdown, and the ten SASS instructions with the most stall samples."""
import csv
import io
import subprocess
import sys

KEYS = ["Grid Size", "Block Size", "launch__cluster_dim_x", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    for rep in sys.argv[1:]:
        raw = ncu_csv(rep, "raw")
        hdr, units = raw[0], raw[1]
        # The source page prints one block per (launch, view) — with --import-source on there are TWO views per launch
        # (SASS and the high-level source), each introduced by a "Kernel Name" row — in launch order.  Pair the blocks
        # with the launches by position and keep, per launch, the view whose "Source" column holds SASS.
        src = ncu_csv(rep, "source")
        views, cur = [], None
        for r in src:
            if r and r[0] == "Kernel Name":
                cur = {"name": r[1] if len(r) > 1 else "", "rows": []}
                views.append(cur)
            elif cur is not None:
                cur["rows"].append(r)
        n_launch = max(len(raw) - 2, 1)
        per = max(len(views) // n_launch, 1)

        def looks_like_sass(view):
            rows = view["rows"]
            if len(rows) < 2 or "Source" not in rows[0]:
                return 0
            i = rows[0].index("Source")
            ops = [r[i].strip().split(" ")[0].lstrip("@!UP0123456789T ") for r in rows[1:40] if len(r) > i]
            return sum(1 for o in ops if o[:3].isupper() or o.split(".")[0].isupper())
        blocks = []
        for k in range(n_launch):
            cand = views[k * per:(k + 1) * per]
            best = max(cand, key=looks_like_sass) if cand else None
            blocks.append(best["rows"] if best else [])
        for k, vals in enumerate(raw[2:]):
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print(f"== {rep.split('/')[-1]}  launch {k}: {d.get('Kernel Name')}")
            base = lambda n: n.split("<")[0].split("(")[0].split("::")[-1].split(" ")[-1]
            if k < len(blocks) and per * k < len(views) and views[per * k]["name"] and base(d.get("Kernel Name", "")) != base(views[per * k]["name"]):
                print(f"   (source view belongs to {views[per * k]['name']!r}: pairing by position failed, SASS lines omitted)")
                blocks[k] = []
            for key in KEYS:
                if key in d and d[key] != "":
                    print(f"{key:75s} {d[key]:>18s} {u.get(key, '')}")
            st = sorted(((float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled")
                         and h.endswith("per_issue_active.ratio") and d[h] not in ("", "n/a")), reverse=True)[:8]
            print("warp stall reasons (cycles per issued instruction): " +
                  ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                            for v, h in st))
            if k < len(blocks) and blocks[k]:
                h2 = blocks[k][0]
                ix = {h: i for i, h in enumerate(h2)}
                rows = blocks[k][1:]
                tot = sum(int(r[ix["# Samples"]]) for r in rows) or 1
                print(f"top stalled SASS instructions (of {tot} samples):")
                for r in sorted(rows, key=lambda r: -int(r[ix["# Samples"]]))[:int(__import__("os").environ.get("NCU_TOP", "10"))]:
                    st2 = {h: int(r[ix[h]]) for h in h2 if h.startswith("stall_") and "(Not" not in h and int(r[ix[h]]) > 0}
                    top = ", ".join(f"{a.replace('stall_', '')} {b}" for a, b in sorted(st2.items(), key=lambda kv: -kv[1])[:2])
                    print(f"   {100.0 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} {top}")
            print()


if __name__ == "__main__":
    main()
