"""Helpers to summarise ncu exports (used to write the notes under profiles/).

    python profiles/ncu_tools.py raw  <raw.csv> [metric-substring ...]
    python profiles/ncu_tools.py src  <source.csv> [top-n]
"""
import csv
import sys

DEFAULT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
           'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'launch__grid_size',
           'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
           'l1tex__t_sector_hit_rate.pct', 'sass__inst_executed_local', 'issue_stalled_long_scoreboard_per',
           'issue_stalled_no_instruction_per', 'issue_stalled_wait_per', 'issue_stalled_math_pipe_throttle_per',
           'issue_stalled_short_scoreboard_per', 'issue_stalled_branch_resolving_per', 'issue_stalled_lg_throttle_per',
           'issue_stalled_not_selected_per', 'issue_stalled_barrier_per', 'smsp__warps_eligible.avg.per_cycle_active']


def raw(path, keys):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    keys = keys or DEFAULT
    for r in rows[2:]:
        print('---', r[hdr.index('Kernel Name')][:90])
        for i, h in enumerate(hdr):
            if any(k in h for k in keys):
                print(f'  {h:80s} {rows[1][i]:>16s} {r[i]}')


def src(path, top=40):
    rows = list(csv.reader(open(path)))
    sections, cur = [], None
    for r in rows:
        if r and r[0] == 'File Path':
            cur = {'file': r[1], 'rows': [], 'hdr': None}
            sections.append(cur)
        elif r and r[0] == 'Function Name':
            cur['func'] = r[1]
        elif r and r[0] == 'Line No':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] is not None and r:
            cur['rows'].append(r)
    allagg = {}
    for s in sections:
        h = s['hdr']
        il, iex, ismp = h.index('Line No'), h.index('Instructions Executed'), h.index('# Samples')
        for r in s['rows']:
            try:
                ln, ex, sm = int(r[il]), int(r[iex] or 0), int(r[ismp] or 0)
            except ValueError:
                continue
            a = allagg.setdefault((s['file'].split('/')[-1], ln), [0, 0, r[1]])
            a[0] += ex
            a[1] += sm
    tot = sum(a[0] for a in allagg.values())
    tots = sum(a[1] for a in allagg.values())
    print(f'total warp instructions {tot}, samples {tots}')
    for (f, ln), a in sorted(allagg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'  {f}:{ln:<5d} {a[0] / tot * 100:5.1f}% instr {a[1] / max(tots, 1) * 100:5.1f}% samples  {a[2][:100]}')


if __name__ == '__main__':
    if sys.argv[1] == 'raw':
        raw(sys.argv[2], sys.argv[3:])
    else:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
