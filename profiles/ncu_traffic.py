"""DRAM bytes per launch of every kernel in an ncu_extract.py CSV -> JSON (bench.py's roofline.traffic):
   python profiles/ncu_traffic.py profiles/r02_ncu_full_step_kernels_final.csv profiles/r02_traffic.json"""
import csv, json, re, sys
rows = {r[0]: r[2:] for r in csv.reader(open(sys.argv[1]))}
names = rows['Kernel Name']
rd, wr = rows['dram__bytes_read.sum'], rows['dram__bytes_write.sum']
units = {r[0]: r[1] for r in csv.reader(open(sys.argv[1]))}
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
sr, sw = scale[units['dram__bytes_read.sum']], scale[units['dram__bytes_write.sum']]
acc = {}
for n, a, b in zip(names, rd, wr):
    k = re.sub(r'^void ', '', n).split('<')[0].split('(')[0]
    acc.setdefault(k, []).append(float(a.replace(',', '')) * sr + float(b.replace(',', '')) * sw)
out = {k: sum(v) / len(v) for k, v in acc.items()}
# bench.py names the stable-sort kernels after the passes they replace
alias = {'eof_tile_hist_kernel': 'eof_cell_hist_kernel', 'eof_tile_scatter_kernel': 'eof_cell_scatter_kernel'}
for a, b in alias.items():
    if a in out: out[b] = out[a]
out['_source'] = sys.argv[1]
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(out, indent=1))
