"""dev tool: refresh profiles/ from ncu captures of one round.

  python tools/make_profiles.py <tag> <launches.csv> <bench.json> C2=<raw.csv> [C3=<raw.csv> ...]

<raw.csv> = `ncu -i rep --page raw --csv` of `ncu --set full ... python tools/stage_time_cfg.py <cfg> 1 2`
(one launch of every kernel of the pass).  Copies them to profiles/<tag>_ncu_full_<cfg>.csv, the launch
list to profiles/<tag>_launches.csv, writes profiles/traffic.json (dram bytes per launch and kernel, keyed by
configuration, stamped with the sha1 of the kernel sources) and prints the markdown tables of the summary.
"""
import collections, csv, json, shutil, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
tag, launches, benchf = sys.argv[1:4]
raws = dict(a.split("=", 1) for a in sys.argv[4:])
names = {'k_frames': 'frames', 'k_vbs': 'frames', 'k_prep': 'prep', 'k_lpc': 'lpc', 'k_search': 'search', 'k_pack': 'pack'}
f = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9}
traffic = {"unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum); one launch = one whole pass of the "
                   "configuration (C2: 38,760 blocks, 158.76 M samples, 635 MB packed s16 in), as in bench.py",
           "source": "profiles/%s_ncu_full_<cfg>.csv (ncu --set full, tools/stage_time_cfg.py <cfg> 1 2)" % tag,
           "csrc_sha1_16": bench.csrc_fingerprint()}
shutil.copy(launches, os.path.join(ROOT, "profiles", "%s_launches.csv" % tag))

def key_of(kn):
    return [v for k, v in names.items() if k in kn][0]

def table(cfgname, raw):
    shutil.copy(raw, os.path.join(ROOT, "profiles", "%s_ncu_full_%s.csv" % (tag, cfgname)))
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = hdr.index
    tr = collections.Counter()
    for r in rows[2:]:
        tr[key_of(r[col('Kernel Name')])] += int(float(r[col('dram__bytes_read.sum')]) * f[units[col('dram__bytes_read.sum')]] +
                                                 float(r[col('dram__bytes_write.sum')]) * f[units[col('dram__bytes_write.sum')]])
    traffic[cfgname] = dict(tr)
    ks = [r[col('Kernel Name')].split('(')[0].replace('void ', '') for r in rows[2:]]
    sel = [i for i, k in enumerate(ks) if not (k.startswith('k_frames') or k.startswith('k_vbs'))]
    def row(name, fmt='%.1f', scale=1.0, label=None):
        if name not in hdr: return
        i = col(name); vals = []
        for j in sel:
            try: vals.append(fmt % (float(rows[2 + j][i]) * scale))
            except ValueError: vals.append(rows[2 + j][i][:10])
        print('| %s | %s |' % (label or name, ' | '.join(vals)))
    print('\n### %s\n' % cfgname)
    print('| metric | ' + ' | '.join(ks[j] for j in sel) + ' |'); print('|---|' + '---|' * len(sel))
    u = units[col('gpu__time_duration.sum')]
    row('gpu__time_duration.sum', '%.0f', {'ms': 1000.0, 'us': 1.0, 's': 1e6, 'ns': 1e-3}.get(u, 1.0), 'duration under ncu, us')
    row('launch__grid_size', '%d', 1, 'grid'); row('launch__block_size', '%d', 1, 'block')
    row('launch__registers_per_thread', '%d', 1, 'registers/thread')
    row('sm__warps_active.avg.pct_of_peak_sustained_active', '%.1f', 1, 'warps active, % of peak')
    row('smsp__inst_executed.sum', '%.0f', 1e-6, 'warp instructions, M')
    row('smsp__issue_active.avg.pct_of_peak_sustained_active', '%.1f', 1, 'issue slots used, %')
    for p_, l_ in (('alu', 'ALU'), ('fma', 'FMA (IMAD)'), ('fmaheavy', 'FMA heavy'), ('fp64', 'FP64'), ('lsu', 'LSU'), ('xu', 'XU')):
        row('sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active' % p_, '%.1f', 1, l_ + ' pipe, %')
    row('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '%.1f', 1, 'DRAM throughput, % of peak')
    for nm, lab in (('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write')):
        i = col(nm); print('| %s, Mbyte | %s |' % (lab, ' | '.join('%.1f' % (float(rows[2 + j][i]) * f[units[i]] / 1e6) for j in sel)))
    row('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', '%.1f', 1e-6, 'shared bank conflicts, M wavefronts')
    row('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', '%.1f', 1e-6, 'shared wavefronts, M')
    print(); print('| stall reason (pc samples) | ' + ' | '.join(ks[j] for j in sel) + ' |'); print('|---|' + '---|' * len(sel))
    for hh in hdr:
        if hh.startswith('smsp__pcsamp_warps_issue_stalled_') and not hh.endswith('not_issued'):
            i = col(hh); vals = [int(rows[2 + j][i] or 0) for j in sel]
            if max(vals) > 1500:
                print('| %s | %s |' % (hh.replace('smsp__pcsamp_warps_issue_stalled_', ''), ' | '.join(str(v) for v in vals)))

for cfgname, raw in raws.items():
    table(cfgname, raw)
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)

b = json.loads(open(benchf).read().strip().splitlines()[-1])
share = b['roofline']['stage_share']
L = list(csv.reader(l for l in open(launches) if l.startswith('"')))
h = L[0]; t = collections.Counter(); c = collections.Counter()
for r in L[1:]:
    kn = r[h.index('Kernel Name')].split('(')[0].replace('void ', '')
    t[kn] += float(r[h.index('Metric Value')]); c[kn] += 1
tot = sum(t.values())
print("\n### launch list\n\n| kernel | launches | total ms | share (ncu) | share (bench.py events, device-resident passes) |\n|---|---|---|---|---|")
for k, v in t.most_common():
    print('| %s | %d | %.2f | %.1f %% | %.1f %% |' % (k, c[k], v / 1e6, 100 * v / tot, 100 * share.get(key_of(k), 0)))
