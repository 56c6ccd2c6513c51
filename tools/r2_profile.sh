#!/bin/bash
# Round-2 measurement pass on one B200 (gpurun): sweeps, ncu captures, launch list, in-graph trace, bench lines.
# Everything lands in gpurun_out/; the summaries are copied into profiles/ by hand afterwards.
O=gpurun_out
mkdir -p $O
T="timeout 240"
$T python tools/agg_sweep.py --folded --cases zinc,molhiv,cifar,pattern --scales 1,16 > $O/r2_sweep_folded.jsonl 2> $O/sweep.err
DGN_TILE=1 $T python tools/agg_sweep.py --folded --cases cifar,pattern --scales 1 > $O/r2_sweep_folded_tile.jsonl 2>> $O/sweep.err
for k in agg_fwd_row agg_bwd_row; do
  $T ncu --set full --clock-control none -k regex:$k --launch-skip 1 -c 2 -f -o /tmp/r2_${k}_x1 \
     python tools/agg_sweep.py --folded --cases zinc --scales 1 > $O/ncu_$k.log 2>&1
  ncu -i /tmp/r2_${k}_x1.ncu-rep --page raw --csv > $O/r2_ncu_${k}_x1_raw.csv 2>/dev/null
  ncu -i /tmp/r2_${k}_x1.ncu-rep --page details --csv > $O/r2_ncu_${k}_x1_details.csv 2>/dev/null
done
$T ncu --set full --clock-control none --import-source on -k regex:'post_fwd|post_bwd|wgrad|pair_gather|norm_pair' --launch-skip 40 -c 10 -f -o /tmp/r2_post_kernels \
   python bench.py --steps 2 --warmup 1 --no-cpu --no-strong > $O/ncu_post.log 2>&1
ncu -i /tmp/r2_post_kernels.ncu-rep --page raw --csv > $O/r2_ncu_post_kernels_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep
# the reports with SASS are ~45 MB each (gpurun_out is capped at 64 MiB): keep the csv pages, and the one report that fits
[ $(stat -c %s /tmp/r2_post_kernels.ncu-rep) -lt 30000000 ] && cp /tmp/r2_post_kernels.ncu-rep $O/
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2_launches_final.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu --no-strong > $O/b_ncu.log 2>&1
$T python tools/step_trace.py --out $O/r2_step_trace.json > $O/r2_step_trace.txt 2>&1
$T python tools/step_trace.py --hidden 45 --aggregators "mean dir1-dx dir1-av" --out $O/r2_step_trace_h45.json > $O/r2_step_trace_h45.txt 2>&1
$T python tools/step_trace.py --hidden 48 --aggregators "mean dir1-dx dir1-av" --out $O/r2_step_trace_h48.json > $O/r2_step_trace_h48.txt 2>&1
DGN_POST_DBG=1 $T python tools/post_phases.py > $O/r2_post_phases.txt 2>&1
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 2> $O/bref.err | tail -1 > $O/r2_bench_reference.json
timeout 400 python bench.py 2> $O/b1.err | tail -1 > $O/r2_bench_1gpu.json
for w in molhiv cifar pattern; do
  $T python bench.py --workload $w --no-cpu --steps 50 --warmup 5 2> $O/b_$w.err | tail -1 > $O/r2_bench_$w.json
done
for h in 45 48; do
  $T python bench.py --hidden $h --aggregators "mean dir1-dx dir1-av" --no-cpu --steps 100 --warmup 10 2>/dev/null | tail -1 > $O/r2_bench_hidden$h.json
done
du -sh $O
python - <<'PY'
import json
for f in ("r2_bench_reference", "r2_bench_1gpu", "r2_bench_molhiv", "r2_bench_cifar", "r2_bench_pattern", "r2_bench_hidden45", "r2_bench_hidden48"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("kind"))
    except Exception as e:
        print(f, "ERR", e)
PY
