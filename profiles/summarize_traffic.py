"""ncu `--set full` captures (gpurun_out/<tag>_<kernel>.ncu-rep) -> profiles/ncu_traffic.json: DRAM bytes per bench scope
for the kernels bench.py reports a roofline for, keyed to the library build they were captured on.

    python profiles/summarize_traffic.py <tag> [workload]

Per kernel: dram__bytes_read.sum + dram__bytes_write.sum of the captured launch x launches per scope (a scope of
bench.py = one layer: two launches for the per-block kernels)."""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else 'c4'
# launches per bench scope (a scope of bench.py = one layer): the tensor kernels run once per filter block
per_scope = {'k_filter_tc': 2, 'k_edge_bwd2_tc': 2, 'k_alpha_aggregate': 1, 'k_qk_bwd_stream': 1, 'k_alpha_bwd': 1}
files = {'k_alpha_aggregate': 'k_alpha_aggregate_scan'}
out = {}
for k, mult in per_scope.items():
    rep = os.path.join(ROOT, 'gpurun_out', f'{tag}_{k}.ncu-rep')
    raw = os.path.join(ROOT, 'profiles', f'{tag}_{files.get(k, k)}_raw.csv')
    if os.path.exists(raw):                      # exported on the GPU box (profiles/r02_ncu.sh)
        txt = open(raw).read()
    elif os.path.exists(rep):
        txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    else:
        continue
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(',', ''))
        u = units[i].lower()
        return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    b = get('dram__bytes_read.sum') + get('dram__bytes_write.sum')
    out[k] = b * mult
    print(k, 'DRAM bytes per launch %.3e' % b, 'duration', vals[hdr.index('gpu__time_duration.sum')], units[hdr.index('gpu__time_duration.sum')])
sys.path.insert(0, ROOT)
from mlff_b200 import _lib
sha = _lib.source_sha16()          # the kernel sources of the captured build (the binary is not bit-reproducible)
json.dump({'tag': tag, 'workload': workload, 'source_sha16': sha, 'bytes_per_scope': out,
           'how': 'ncu --set full --clock-control none, one launch per kernel, x launches per bench scope'},
          open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'), 'w'), indent=1)
print('wrote profiles/ncu_traffic.json for sources', sha)
