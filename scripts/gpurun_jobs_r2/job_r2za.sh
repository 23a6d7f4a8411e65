set -x
mkdir -p gpurun_out
for o in 0 1 2; do
  GPK_DIAG_OVL=$o timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:potrf_diag -c 32 --csv --log-file gpurun_out/diag_dur_ovl$o.csv python scripts/one_eval.py 4096 1 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/diag_dur_ovl$o.csv') if not l.startswith('=='))]
v=[float(r['Metric Value']) for r in rows if r['Metric Name']=='gpu__time_duration.sum']
u=rows[0]['Metric Unit'] if rows else '?'
print('ovl$o', len(v), 'launches, mean', sum(v)/max(1,len(v)), 'min', min(v), 'max', max(v), u)
PY
done
for v in 1 2 1 2; do
  GPK_DIAG_OVL=$v timeout 100 python scripts/quick_eval.py 16384 12 "ovl$v" | tail -1
done
for v in 1 2 1 2; do
  GPK_DIAG_OVL=$v timeout 100 python scripts/quick_eval.py 6144 16 "ovl$v" | tail -1
done
