#!/bin/bash
# One GPU call: the ncu launch list of the default command, one `ncu --set full` capture per workload and
# recurrence form, the application-replay traffic of the constant-bank sweep, then the default bench line and
# the reference arm.  The .ncu-rep files are
# summarised ON THE BOX (gpurun brings back at most 64 MiB): only the markdown summaries and
# traffic.json return, under gpurun_out/profiles/; copy them into profiles/ afterwards.
set -x
O=gpurun_out/profiles
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file /tmp/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench_all.log 2>&1
python scripts/summarize_profiles.py r2_default --out $O --launches /tmp/r2_launches.csv
cp profiles/traffic.json $O/traffic.json 2>/dev/null
while read -r w f pat key per skip; do
  ncu --set full --clock-control none --import-source on -k "regex:$pat" --launch-skip ${skip:-0} -c 4 -o /tmp/r2_${w}_f$f python scripts/ncu_target.py $w $f 1 > gpurun_out/r2_ncu_${w}_f$f.log 2>&1
  python scripts/summarize_profiles.py r2_${w}_f$f --out $O --rep /tmp/r2_${w}_f$f.ncu-rep --traffic $key --kernel ${pat%%|*} --per-step $per
  rm -f /tmp/r2_${w}_f$f.ncu-rep
done <<SPECS
c2 0 numerov_sweep c2 1
c2 1 numerov_sweep c2_dform 1
c5 0 numerov_cbank c5 715 10
c5 1 numerov_cbank c5_dform 715 10
c3 0 numerov_sweep|segment_combine c3 1
c3 1 numerov_sweep|segment_combine c3_dform 1
c4 0 numerov_sweep c4 1
c4 1 numerov_sweep c4_dform 1
cooley 1 cooley cooley 1
SPECS
# c5 traffic: kernel replay flushes L2 before every pass, so the figures above always show the carried state
# coming from DRAM; application replay without cache control sees the L2 a real sweep leaves behind
for f in 0 1; do
  ncu --replay-mode application --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:numerov_cbank \
      --launch-skip 120 -c 120 --csv --log-file /tmp/r2_c5_f${f}_appreplay.csv python scripts/ncu_target.py c5 $f 1 > gpurun_out/r2_ncu_c5_f${f}_app.log 2>&1
  per=$(grep -o "cbank launches per sweep [0-9]*" gpurun_out/r2_ncu_c5_f${f}_app.log | tail -1 | grep -o "[0-9]*$")
  key=c5; [ $f = 1 ] && key=c5_dform
  python scripts/summarize_profiles.py r2_c5_f${f}_appreplay --out $O --traffic-csv /tmp/r2_c5_f${f}_appreplay.csv --traffic $key --per-step ${per:-715}
done
# the default bench line and the reference arm LAST, with the traffic figures captured above (same kernel sources)
cp $O/traffic.json profiles/traffic.json
python bench.py > $O/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference > $O/r2_bench_ref.json 2>> gpurun_out/r2_bench_n1.err
ls -la $O
