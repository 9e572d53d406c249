"""dev tool: refresh profiles/ from a capture.

  python tools/make_profiles.py <raw.csv from `ncu --page raw --csv`> <launches.csv> <bench.json>
writes profiles/r01_ncu_full_final.csv, r01_launches_final.csv, traffic.json and prints the
markdown tables for profiles/r01_summary.md.
"""
import collections, csv, json, shutil, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, launches, bench = sys.argv[1:4]
shutil.copy(raw, os.path.join(ROOT, "profiles", "r01_ncu_full_final.csv"))
shutil.copy(launches, os.path.join(ROOT, "profiles", "r01_launches_final.csv"))
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = hdr.index
names = {'k_frames': 'frames', 'k_prep': 'prep', 'k_lpc': 'lpc', 'k_search': 'search', 'k_pack': 'pack'}
f = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9}
traffic = {}
for r in rows[2:]:
    key = [v for k, v in names.items() if k in r[col('Kernel Name')]][0]
    traffic[key] = int(float(r[col('dram__bytes_read.sum')]) * f[units[col('dram__bytes_read.sum')]] +
                       float(r[col('dram__bytes_write.sum')]) * f[units[col('dram__bytes_write.sum')]])
json.dump({"unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum); one launch = the whole C2 stream "
                   "(38,760 blocks, 158.76 M samples, 635 MB packed s16 in), as in bench.py",
           "source": "profiles/r01_ncu_full_final.csv (ncu --set full, tools/stage_time.py 8 3600 1)", **traffic},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
b = json.load(open(bench))
share = b['roofline']['stage_share']
L = list(csv.reader(l for l in open(launches) if l.startswith('"')))
h = L[0]; t = collections.Counter(); c = collections.Counter()
for r in L[1:]:
    kn = r[h.index('Kernel Name')].split('(')[0].replace('void ', '')
    t[kn] += float(r[h.index('Metric Value')]); c[kn] += 1
tot = sum(t.values())
print("| kernel | launches | total ms | share (ncu) | share (bench.py events, device-resident passes) |\n|---|---|---|---|---|")
for k, v in t.most_common():
    key = [vv for kk, vv in names.items() if kk in k][0]
    print('| %s | %d | %.2f | %.1f %% | %.1f %% |' % (k, c[k], v / 1e6, 100 * v / tot, 100 * share.get(key, 0)))
print()
ks = [r[col('Kernel Name')].split('(')[0].replace('void ', '') for r in rows[2:]]
sel = [i for i, k in enumerate(ks) if not k.startswith('k_frames')]
def row(name, fmt='%.1f', scale=1.0, label=None):
    i = col(name); vals = []
    for j in sel:
        try: vals.append(fmt % (float(rows[2 + j][i]) * scale))
        except ValueError: vals.append(rows[2 + j][i][:10])
    print('| %s | %s |' % (label or name, ' | '.join(vals)))
print('| metric | ' + ' | '.join(ks[j] for j in sel) + ' |'); print('|---|' + '---|' * len(sel))
u = units[col('gpu__time_duration.sum')]
row('gpu__time_duration.sum', '%.0f', {'ms': 1000.0, 'us': 1.0, 's': 1e6}.get(u, 1.0), 'duration under ncu, us')
row('launch__grid_size', '%d', 1, 'grid'); row('launch__block_size', '%d', 1, 'block')
row('launch__registers_per_thread', '%d', 1, 'registers/thread')
row('sm__warps_active.avg.pct_of_peak_sustained_active', '%.1f', 1, 'warps active, % of peak')
row('smsp__inst_executed.sum', '%.0f', 1e-6, 'warp instructions, M')
row('smsp__issue_active.avg.pct_of_peak_sustained_active', '%.1f', 1, 'issue slots used, %')
for p_, l_ in (('alu', 'ALU'), ('fma', 'FMA (IMAD)'), ('fp64', 'FP64'), ('lsu', 'LSU'), ('xu', 'XU')):
    row('sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active' % p_, '%.1f', 1, l_ + ' pipe, %')
row('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '%.1f', 1, 'DRAM throughput, % of peak')
for nm, lab in (('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write')):
    i = col(nm); print('| %s, %s | %s |' % (lab, units[i], ' | '.join('%.2f' % float(rows[2 + j][i]) for j in sel)))
row('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', '%.1f', 1e-6, 'shared bank conflicts, M wavefronts')
row('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', '%.1f', 1e-6, 'shared wavefronts, M')
print(); print('| stall reason (pc samples) | ' + ' | '.join(ks[j] for j in sel) + ' |'); print('|---|' + '---|' * len(sel))
for hh in hdr:
    if hh.startswith('smsp__pcsamp_warps_issue_stalled_') and not hh.endswith('not_issued'):
        i = col(hh); vals = [int(rows[2 + j][i] or 0) for j in sel]
        if max(vals) > 1500:
            print('| %s | %s |' % (hh.replace('smsp__pcsamp_warps_issue_stalled_', ''), ' | '.join(str(v) for v in vals)))
