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
    # keep the LAST solve only: it starts at the last reset_guess_kernel (set-up and the warm-up solve come before it)
    order = list(per.values())
    starts = [i for i, d in enumerate(order) if d['name'].startswith('reset_guess_kernel')]
    if starts:
        order = order[starts[-1]:]
        ends = [i for i, d in enumerate(order) if d['name'].startswith('cg_outer_kernel')]     # the solve ends with its last true-residual stage
        if ends:
            order = order[:ends[-1] + 1]
    agg = collections.OrderedDict()
    for d in order:
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


def full(path, match=None):
    """metrics of the first captured launch whose kernel name contains `match`"""
    rows = list(csv.reader(open(path)))
    units = dict(zip(rows[0], rows[1]))
    for r in rows[2:]:
        d = dict(zip(rows[0], r))
        if match is None or match in d.get("Kernel Name", ""):
            return d, units
    raise KeyError(match)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2y"          # tools/gpu_prof3.sh <tag>
    if os.path.isfile(os.path.join(G, tag + "_launches.csv")):
        launch_list(os.path.join(G, tag + "_launches.csv"), os.path.join(OUT, "launches_16384_k3_fp32_summary.txt"))
        import shutil
        shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(OUT, "launches_16384_k3_fp32.csv"))
    caps = [("x/r update+check", "cg_update_tile_kernel<float>", tag + "_vec_raw.csv", "cg_update_tile_kernel"),
            ("fine down", "sl_down_kernel<float, 0, 1> (lean chunks)", tag + "_sl_raw.csv", "sl_down_kernel<float, 0, 1>"),
            ("fine up", "sl_up_kernel<float, 0, 1> (lean chunks)", tag + "_sl_raw.csv", "sl_up_kernel<float, 0, 1>"),
            ("stencil+alpha", "sl_lap_kernel<float, float, 0, 0> (lean chunks)", tag + "_sl_raw.csv", "sl_lap_kernel<float, float, 0, 0>"),
            ("outer residual", "sl_lap_kernel<double, float, 0, 1> (fp64 true residual, lean chunks)", tag + "_sl_raw.csv", "sl_lap_kernel<double, float, 0, 1>"),
            ("level-1 down", "cs_leg_kernel<float, 0> level 1 (8192^2)", tag + "_csdown_raw.csv", "cs_leg_kernel<float, 0>"),
            ("level-1 up", "cs_leg_kernel<float, 1> level 1 (8192^2)", tag + "_csup_raw.csv", "cs_leg_kernel<float, 1>"),
            ("beta+p update", "cg_pupdate_tile_kernel<float>", tag + "_vec_raw.csv", "cg_pupdate_tile_kernel")]
    traffic = {}
    with open(os.path.join(OUT, "ncu_full_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["phase", "kernel", "Kernel Name (ncu)"] + KEYS)
        for phase, label, fn, match in caps:
            p = os.path.join(G, fn)
            if not os.path.isfile(p) or len(list(csv.reader(open(p)))) < 3:
                continue
            d, units = full(p, match)
            w.writerow([phase, label, d.get("Kernel Name", "")] + [d.get(k, "") + " " + units.get(k, "") for k in KEYS])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "GB": 1e9, "MB": 1e6, "KB": 1e3, "B": 1.0}
            rd = float(d["dram__bytes_read.sum"]) * scale.get(units["dram__bytes_read.sum"], 1.0)
            wr = float(d["dram__bytes_write.sum"]) * scale.get(units["dram__bytes_write.sum"], 1.0)
            traffic[phase] = {"kernel": label, "dram_read_bytes": rd, "dram_write_bytes": wr,
                              "source": "profiles/r2/ncu_full_summary.csv (ncu --set full --clock-control none, python tools/profile_step.py 16384 1, workload of bench.py at N=1)"}
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(open(os.path.join(OUT, "launches_16384_k3_fp32_summary.txt")).read())
    print(json.dumps({k: (v["dram_read_bytes"] + v["dram_write_bytes"]) / 1e9 for k, v in traffic.items()}, indent=1))


if __name__ == "__main__":
    main()
