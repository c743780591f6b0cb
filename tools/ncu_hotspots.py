#!/usr/bin/env python3
"""Per-source-line hotspots of one kernel launch in an ncu report (needs -lineinfo and --import-source on).

  python tools/ncu_hotspots.py report.ncu-rep [--launch K] [--top N]

Reads `ncu -i report --page source --csv --print-source sass,cuda` and prints, per source line: share of stall samples,
share of executed warp instructions, active threads per instruction, and the local-memory (stack / spill) instructions
that line issued -- plus per-file-function totals.  Used to write the summaries under profiles/."""
import argparse
import collections
import csv
import io
import subprocess
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--launch', type=int, default=0)
    ap.add_argument('--top', type=int, default=25)
    ap.add_argument('--kernel', default='regex:render_pass')
    a = ap.parse_args()
    out = subprocess.run(['ncu', '-i', a.report, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name', a.kernel,
                          '--launch-skip', str(a.launch), '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr = None, None
    lines = {}      # (file, line) -> dict
    cur = None
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if len(r) == 2 and r[0] == 'Function Name':
            continue
        if r and r[0] == 'Line No':
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            # two columns are called "Source": the first is the CUDA line, the second the SASS text
            src_cols = [i for i, h in enumerate(hdr) if h == 'Source']
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        def num(name):
            try:
                return float(r[ix[name]])
            except (ValueError, KeyError):
                return 0.0
        if r[0] != '':
            cur = (fname, int(r[0]))
            lines[cur] = {'text': r[src_cols[0]].strip(), 'samples': num('# Samples'), 'inst': num('Instructions Executed'),
                          'tinst': num('Thread Instructions Executed'), 'local': 0.0, 'stalls': collections.Counter()}
            for h in hdr:
                if h.startswith('stall_') and '(' not in h:
                    lines[cur]['stalls'][h] += num(h)
        elif cur is not None:
            if r[ix['Address Space']] == 'Local':
                lines[cur]['local'] += num('Instructions Executed')
    tot_s = sum(v['samples'] for v in lines.values()) or 1
    tot_i = sum(v['inst'] for v in lines.values()) or 1
    tot_l = sum(v['local'] for v in lines.values()) or 1
    print('total stall samples %d, warp instructions %d, of which local-memory ld/st %d (%.1f %%)' % (tot_s, tot_i, tot_l, 100 * tot_l / tot_i))
    st = collections.Counter()
    for v in lines.values():
        st.update(v['stalls'])
    print('stall reasons (share of samples): ' + ', '.join('%s %.1f%%' % (k[6:], 100 * n / tot_s) for k, n in st.most_common(8)))

    def show(title, key):
        print('--- top by ' + title)
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][key])[:a.top]:
            print('%-16s %5d  %5.1f%% smp  %5.1f%% ins  %5.1f%% local  thr/inst %4.1f | %s' % (
                f, ln, 100 * v['samples'] / tot_s, 100 * v['inst'] / tot_i, 100 * v['local'] / tot_l,
                v['tinst'] / v['inst'] if v['inst'] else 0, v['text'][:110]))
    show('stall samples', 'samples')
    show('warp instructions', 'inst')
    show('local-memory instructions', 'local')


if __name__ == '__main__':
    main()
