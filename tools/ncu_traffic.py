#!/usr/bin/env python3
"""DRAM traffic of the dominant kernel of a bench config from an `ncu --set full` capture, for bench.py's roofline.traffic.

  python tools/ncu_traffic.py <config> <report.ncu-rep> [--source profiles/<file>]

Sums dram__bytes_read.sum + dram__bytes_write.sum over the captured launches of render_pass_kernel (the capture holds the
launches of ONE step: 5 for config 4 -- primary pass + 4 bounce passes --, 1 for the single-pass configs) and records
the total in profiles/r02_traffic.json next to the kernel name and the file the numbers come from."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def main():
    config, rep = sys.argv[1], sys.argv[2]
    source = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == '--source' else os.path.relpath(rep, ROOT)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    total, launches, kernel, ms = 0.0, 0, None, 0.0
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        if 'render_pass_kernel' not in name:
            continue
        kernel = name.split('(')[0].replace('void ', '').replace('ntr::', '').replace(' ', '')
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(m)
            total += float(r[i]) * UNIT[units[i]]
        i = hdr.index('gpu__time_duration.sum')
        ms += float(r[i]) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[units[i]]
        launches += 1
    path = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[config] = {'kernel': kernel, 'bytes_per_launch': total, 'launches': launches, 'ms_under_ncu': ms,
                    'what': 'dram__bytes_read.sum + dram__bytes_write.sum summed over the %d launch(es) of the kernel in one step '
                            '(ncu --set full --clock-control none)' % launches,
                    'source': source}
    json.dump(data, open(path, 'w'), indent=1)
    print(config, data[config])


if __name__ == '__main__':
    main()
