#!/bin/bash
# Round-end measurement on one B200.  Part "a": full GPU test suite, smoke, bench lines (both arms, both detectors, single
# expert), per-launch GEMM timings, ncu launch lists.  Part "b": two `ncu --set full` captures, summarised on the box
# (the reports themselves exceed what gpurun copies back).  Outputs: gpurun_out/final_*.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
if [ "$1" != "b" ]; then
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/final_smi.txt 2>&1
python -m pytest tests -q -m gpu -p no:cacheprovider > $O/final_gpu_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/final_gpu_tests.log)"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/final_smoke.log)"
python bench.py > $O/final_bench_n1.json 2> $O/final_bench_n1.err; echo "bench rc=$? $(cut -c1-160 $O/final_bench_n1.json)"
python bench.py --impl reference > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "ref rc=$? $(cut -c1-200 $O/final_bench_reference.json)"
python bench.py --arch neutron --no-cpu-baseline > $O/final_bench_neutron_n1.json 2>/dev/null; echo "neutron $(cut -c1-160 $O/final_bench_neutron_n1.json)"
python bench.py --experts 1 --no-cpu-baseline > $O/final_bench_c2_e1.json 2>/dev/null; echo "E=1 $(cut -c1-160 $O/final_bench_c2_e1.json)"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --per-launch 2>&1 >/dev/null | grep "es_" > $O/final_gemm_per_launch_proton.txt
python bench.py --arch neutron --steps 5 --warmup 3 --no-cpu-baseline --per-launch 2>&1 >/dev/null | grep "es_" > $O/final_gemm_per_launch_neutron.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launches_train_step.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --ncu-step 1 > $O/ncu1.log 2>&1; echo "launch list rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launches_neutron_train_step.csv \
    python bench.py --arch neutron --steps 2 --warmup 1 --no-cpu-baseline --ncu-step 1 > $O/ncu1n.log 2>&1; echo "launch list (neutron) rc=$?"
else
ncu --set full --clock-control none --profile-from-start off -k regex:igemm -c 20 -f -o /tmp/final_prof_igemm \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --ncu-step 1 > $O/ncu2.log 2>&1; echo "ncu igemm rc=$?"
python tools/ncu_summary.py /tmp/final_prof_igemm.ncu-rep $O/final_ncu_igemm_summary.csv; mv $O/roofline_traffic.json $O/final_roofline_traffic.json 2>/dev/null
ncu --set full --clock-control none --profile-from-start off -k 'regex:loss|router|hinge|disc_|gn_|adam|gen_out|ln_' -c 24 -f -o /tmp/final_prof_hbm \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --ncu-step 1 > $O/ncu3.log 2>&1; echo "ncu hbm rc=$?"
python tools/ncu_summary.py /tmp/final_prof_hbm.ncu-rep $O/final_ncu_hbm_summary.csv; rm -f $O/roofline_traffic.json
ls -la /tmp/final_prof_*.ncu-rep
fi
ls -la $O/final_* | awk '{print $5, $9}'
