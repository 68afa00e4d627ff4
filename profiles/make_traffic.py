"""profiles/traffic.json from an `ncu --set full` report of probe_size.py 2048 (one launch per kernel class, all instances active).

usage: python profiles/make_traffic.py <report>.ncu-rep <bench line .json of the same build> "<note>"
The algorithmic bytes of a launch = cells of that launch x mseetc_bytes_per_cell (the library's own table, taken here from the
`roofline.kernels` block that bench.py printed for the same build)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

rep, benchline, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
H = rows[0]
col = lambda r, name: r[H.index(name)]
unit = lambda name: rows[1][H.index(name)]
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3}
CLASS = {'k_cell_step': 'cell_step', 'k_inst_alpha': 'inst_alpha', 'k_cell_trial_eval': 'cell_trial', 'k_inst_kkt<1>': 'inst_decide',
         'k_step_pit': 'inst_step', 'k_step<': 'inst_step'}
INSTANCES, CELLS = 2048, 2048 * 301
bpc = {k: v['bytes_per_cell'] for k, v in json.load(open(benchline))['roofline']['kernels'].items()}
out = {'note': note, 'captures': {}}
for r in rows[2:]:
    name = col(r, 'Kernel Name')
    cls = next((v for k, v in CLASS.items() if k in name), None)
    if cls is None:
        continue
    dram = (float(col(r, 'dram__bytes_read.sum')) * scale[unit('dram__bytes_read.sum')] +
            float(col(r, 'dram__bytes_write.sum')) * scale[unit('dram__bytes_write.sum')])
    us = float(col(r, 'gpu__time_duration.sum')) * scale[unit('gpu__time_duration.sum')]
    alg = CELLS * bpc[cls]
    out['captures'][cls] = dict(kernel=name.split('(')[0], dram_bytes_per_launch=dram, algorithmic_bytes_per_launch=alg, ratio=round(dram / alg, 3),
                                duration_us=us, grid=float(col(r, 'launch__grid_size')), hbm_gbs_at_capture=round(dram / us / 1e3, 1))
    out[cls] = dram
json.dump(out, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
print(json.dumps(out['captures'], indent=1))
