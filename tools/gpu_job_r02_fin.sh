# round 2, closing job: whole GPU suite, smoke, bench (configs 1-3, sequential schedule, reference arm), ncu launch lists, live shares
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02fin_test_all.log 2>&1; tail -3 gpurun_out/r02fin_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02fin_test_all.log | cut -c1-300 | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fin_smoke.log 2>&1; tail -2 gpurun_out/r02fin_smoke.log | cut -c1-300
for c in 1 2 3; do timeout 900 python bench.py --steps 20 --warmup 3 --config $c > gpurun_out/r02fin_bench_cfg$c.json 2>gpurun_out/r02fin_bench_cfg$c.err; cut -c1-200 gpurun_out/r02fin_bench_cfg$c.json; tail -2 gpurun_out/r02fin_bench_cfg$c.err; done
TTDG_OVERLAP=0 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02fin_bench_cfg1_sequential.json 2>gpurun_out/r02fin_bench_cfg1_sequential.err; cut -c1-200 gpurun_out/r02fin_bench_cfg1_sequential.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02fin_bench_reference.json 2>gpurun_out/r02fin_bench_reference.err; cut -c1-200 gpurun_out/r02fin_bench_reference.json
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/r02fin_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02fin_ncu_bench.log 2>&1; tail -1 gpurun_out/r02fin_ncu_bench.log | cut -c1-120; wc -l gpurun_out/r02fin_launches_bench.csv
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02fin_launches_bench.csv', errors='ignore')) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(',', ''))
    except ValueError:
        continue
    name = r[ik].split('(')[0][:90]
    tot[name] += v / 1e3; cnt[name] += 1
s = sum(tot.values())
with open('gpurun_out/r02fin_launches_bench_summary.csv', 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 python bench.py --steps 2 --warmup 3 (first 6000 launches: warm-up + timed + e2e steps; cold-cache, serialised)\n')
    f.write('kernel,launches,total_us,share_pct\n')
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:50]:
        f.write('%s,%d,%.1f,%.2f\n' % (k, cnt[k], v, 100 * v / s))
print(open('gpurun_out/r02fin_launches_bench_summary.csv').read()[:1500])
PY
rm -f gpurun_out/r02fin_launches_bench.csv
timeout 600 $NCU --profile-from-start off --set full --import-source on -k regex:gagm_kernel -c 1 -o /tmp/prof_gagm -f python tools/run_kernels.py full_step 3 > gpurun_out/r02fin_ncu_gagm.log 2>&1; tail -1 gpurun_out/r02fin_ncu_gagm.log | cut -c1-200
ncu -i /tmp/prof_gagm.ncu-rep --page raw --csv > gpurun_out/r02fin_gagm_ncu_full_raw.csv 2>/dev/null
ncu -i /tmp/prof_gagm.ncu-rep --page source --csv > gpurun_out/r02fin_gagm_ncu_source.csv 2>/dev/null; wc -c gpurun_out/r02fin_gagm_ncu_source.csv
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/r02fin_busy_sequential.csv 2>/dev/null; head -2 gpurun_out/r02fin_busy_sequential.csv | cut -c1-160
timeout 300 python tools/run_kernels.py busy 3 gaps overlap > gpurun_out/r02fin_busy_overlap.csv 2>/dev/null; head -2 gpurun_out/r02fin_busy_overlap.csv | cut -c1-160
timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02fin_layers_fp32.csv 2>/dev/null; head -1 gpurun_out/r02fin_layers_fp32.csv
timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-800 > gpurun_out/r02fin_gagm_bench.txt; head -2 gpurun_out/r02fin_gagm_bench.txt | cut -c1-400
du -sh gpurun_out
