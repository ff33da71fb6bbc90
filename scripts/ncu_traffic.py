"""Sum DRAM bytes and durations over the launches of one MatMult in an .ncu-rep and write
profiles/traffic.json + a text summary.  usage: ncu_traffic.py rep.ncu-rep KEY launches_per_matmult out.txt"""
import csv
import io
import json
import os
import subprocess
import sys

rep, key, per, out_txt = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]


def col(name):
    i = hdr.index(name)
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0}.get(units[i], 1.0)
    return [float(r[i].replace(',', '')) * scale for r in rows[2:]]


rd, wr, dur = col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), col('gpu__time_duration.sum')
names = [r[hdr.index('Kernel Name')] for r in rows[2:]]
n = (len(rd) // per) * per
tot = sum(rd[:n]) + sum(wr[:n])
per_mm = tot / (n // per)
lines = [f'report: {os.path.basename(rep)}   launches captured: {len(rd)}   launches per MatMult: {per}']
for i in range(len(rd)):
    lines.append(f'  launch {i}: {names[i][:70]}  {dur[i]*1e3:8.3f} ms  dram read {rd[i]/1e9:7.2f} GB  write {wr[i]/1e9:7.2f} GB  '
                 f'-> {(rd[i]+wr[i])/dur[i]/1e9:7.0f} GB/s')
lines.append(f'DRAM traffic per MatMult: {per_mm/1e9:.2f} GB ; summed kernel time per MatMult {sum(dur[:n])/(n//per)*1e3:.3f} ms (under ncu, cold, serialised)')
open(out_txt, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
tpath = os.path.join(os.path.dirname(out_txt), 'traffic.json')
data = json.load(open(tpath)) if os.path.exists(tpath) else {}
data[key] = per_mm
json.dump(data, open(tpath, 'w'), indent=1)
