#!/usr/bin/env python
"""Summarise an ncu --set full report (raw page CSV) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv, sys, subprocess, io, json, os
_pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
PEAK, PEAK_SRC = (json.load(open(_pk))["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if os.path.exists(_pk) else (6650.0, "fallback")
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if name in seen and '--all' not in sys.argv:
        continue
    seen.add(name)
    print('----')
    try:
        def val(k):
            i = hdr.index(k); v = float(r[i].replace(',', '')); u = units[i]
            scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1.0, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0}.get(u, 1.0)
            return v * scale
        by = val('dram__bytes_read.sum') + val('dram__bytes_write.sum'); t = val('gpu__time_duration.sum')
        print(f"{'ACHIEVED HBM (dram bytes / duration)':80s} {by / t / 1e9:.1f} GB/s = {by / t / 1e9 / PEAK:.3f} of the {PEAK:.1f} GB/s {PEAK_SRC} peak ({by / 1e6:.1f} MB in {t * 1e6:.1f} us)")
    except Exception as e:
        print('achieved HBM: n/a', e)
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:80s} {r[i]} {units[i]}")
