"""Turns the ncu exports in gpurun_out/ into the tracked summaries under profiles/r2/ (and profiles/traffic.json):
  * launches_16384_k3_fp32_summary.txt  -- launch list of one solve, aggregated by kernel (time share, DRAM bytes per launch)
  * ncu_full_summary.csv                -- `ncu --set full` key metrics of the dominant kernels
  * ../traffic.json                     -- DRAM bytes per launch per phase of bench.py's kernel table"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles", "r2")
os.makedirs(OUT, exist_ok=True)


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    kn, mn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    gs = hdr.index('Grid Size')
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r'\(hg::.*', '', r[kn]).replace('void hg::', '').replace('void ', '')
        name = re.sub(r'\(.*', '', name)
        d = per.setdefault(r[idc], {'name': name, 'grid': r[gs]})
        d[r[mn]] = float(r[mv].replace(',', ''))
    agg = collections.OrderedDict()
    for d in per.values():
        key = d['name'] + ('  grid ' + d['grid'] if ('cs_leg' in d['name'] or 'cdown' in d['name'] or 'cup_' in d['name']) else '')
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0) / 1e6; a[2] += d.get('dram__bytes_read.sum', 0) / 1e9; a[3] += d.get('dram__bytes_write.sum', 0) / 1e9
    tot = sum(a[1] for a in agg.values())
    with open(out, 'w') as f:
        f.write("# one device-resident solve of the headline workload (16384^2 RGB, 50 % unknown in 64x64 tiles, fp32 storage, rtol 1e-6, 7 iterations)\n")
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (second solve of tools/profile_step.py 16384 2)\n")
        f.write("# per-launch times under ncu are serialised and cold-cache: compare SHARES with bench.py's `kernels` table, not absolutes\n")
        f.write("total %.2f ms over %d launches\n" % (tot, sum(a[0] for a in agg.values())))
        f.write("%-58s %4s %9s %6s %9s %8s %8s %6s\n" % ("kernel", "n", "total ms", "share", "avg ms", "rd GB/l", "wr GB/l", "TB/s"))
        for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda x: -x[1][1]):
            if t / tot < 0.001:
                continue
            f.write("%-58s %4d %9.2f %5.1f%% %9.3f %8.3f %8.3f %6.2f\n" % (k[:58], n, t, 100 * t / tot, t / n, rd / n, wr / n, (rd + wr) / t if t else 0))
    return agg


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']


def full(path):
    rows = list(csv.reader(open(path)))
    d = dict(zip(rows[0], rows[2]))
    units = dict(zip(rows[0], rows[1]))
    return d, units


def main():
    if os.path.isfile(os.path.join(G, "r2s_launches.csv")):
        launch_list(os.path.join(G, "r2s_launches.csv"), os.path.join(OUT, "launches_16384_k3_fp32_summary.txt"))
    caps = [("x/r update+check", "cg_update_tile_kernel<float>", "r2o_cg_update_tile_kernel_raw.csv"),
            ("fine down", "sl_down_kernel<float, 0, 1> (lean chunks)", "r2o_sl_down_kernel_raw.csv"),
            ("fine up", "sl_up_kernel<float, 0, 1> (lean chunks)", "r2o_sl_up_kernel_raw.csv"),
            ("stencil+alpha", "sl_lap_kernel<float, float, 0, 0> (lean chunks)", "r2o_sl_lap_kernel_raw.csv"),
            ("level-1 down", "cs_leg_kernel<float, 0> level 1 (8192^2)", "r2s_cs_leg_kernel_raw.csv"),
            ("level-1 up", "cs_leg_kernel<float, 1> level 1 (8192^2)", "r2s_cs_up_raw.csv"),
            ("beta+p update", "cg_pupdate_tile_kernel<float>", "r2s_cg_pupdate_tile_kernel_raw.csv")]
    traffic = {}
    with open(os.path.join(OUT, "ncu_full_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["phase", "kernel", "Kernel Name (ncu)"] + KEYS)
        for phase, label, fn in caps:
            p = os.path.join(G, fn)
            if not os.path.isfile(p) or len(list(csv.reader(open(p)))) < 3:
                continue
            d, units = full(p)
            w.writerow([phase, label, d.get("Kernel Name", "")] + [d.get(k, "") + " " + units.get(k, "") for k in KEYS])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(d["dram__bytes_read.sum"]) * scale.get(units["dram__bytes_read.sum"], 1.0)
            wr = float(d["dram__bytes_write.sum"]) * scale.get(units["dram__bytes_write.sum"], 1.0)
            traffic[phase] = {"kernel": label, "dram_read_bytes": rd, "dram_write_bytes": wr,
                              "source": "profiles/r2/ncu_full_summary.csv (ncu --set full --clock-control none, python tools/profile_step.py 16384 1, workload of bench.py at N=1)"}
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(open(os.path.join(OUT, "launches_16384_k3_fp32_summary.txt")).read())
    print(json.dumps({k: (v["dram_read_bytes"] + v["dram_write_bytes"]) / 1e9 for k, v in traffic.items()}, indent=1))


if __name__ == "__main__":
    main()
