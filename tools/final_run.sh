#!/bin/bash
# Round-end measurement on one B200 (outputs: gpurun_out/final_*; copy what is to be judged into profiles/r02_*).
#   a   full GPU test suite, smoke, bench lines (both arms, CUDA-graph variant), per-launch GEMM timings
#   b   ncu launch lists of one train step (proton, neutron)                         [one ncu use per gpurun call]
#   c   `ncu --set full` of the tensor-core family, summarised on the box
#   d   `ncu --set full` of the HBM-bound kernels (loss tails, gating, norms, Adam), summarised on the box
#   e   accuracy equivalence over a 300-step trajectory (reference fp32 eager on the GPU vs this build)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu-baseline --no-hbm-kernels --no-extra"
case "$1" in
a)
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/final_smi.txt 2>&1
python -m pytest tests -q -m gpu -p no:cacheprovider > $O/final_gpu_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/final_gpu_tests.log)"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/final_smoke.log)"
python bench.py --steps 20 --warmup 5 > $O/final_bench_n1.json 2> $O/final_bench_n1.err; echo "bench rc=$? $(cut -c1-160 $O/final_bench_n1.json)"
python bench.py --impl reference --steps 5 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "ref rc=$? $(cut -c1-200 $O/final_bench_reference.json)"
python bench.py --cuda-graph $B > $O/final_bench_graph.json 2>/dev/null; echo "graph $(cut -c1-160 $O/final_bench_graph.json)"
python bench.py --arch neutron --cuda-graph $B > $O/final_bench_neutron_graph.json 2>/dev/null; echo "neutron graph $(cut -c1-160 $O/final_bench_neutron_graph.json)"
python bench.py --steps 5 --warmup 3 $B --per-launch 2>&1 >/dev/null | grep "es_" > $O/final_gemm_per_launch_proton.txt
python bench.py --arch neutron --steps 5 --warmup 3 $B --per-launch 2>&1 >/dev/null | grep "es_" > $O/final_gemm_per_launch_neutron.txt
python tools/inference_sweep.py --showers 2000000 > $O/final_inference_sweep_2M.json 2>/dev/null; echo "sweep $(cut -c1-200 $O/final_inference_sweep_2M.json)"
;;
b)
python bench.py --steps 2 --warmup 3 $B > $O/plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launches_train_step.csv \
    python bench.py --steps 2 --warmup 3 $B --ncu-step 1 > $O/ncu1.log 2>&1; echo "launch list rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launches_neutron_train_step.csv \
    python bench.py --arch neutron --steps 2 --warmup 3 $B --ncu-step 1 > $O/ncu1n.log 2>&1; echo "launch list (neutron) rc=$?"
;;
c)
python bench.py --steps 2 --warmup 3 $B > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -k regex:igemm -c 72 -f -o /tmp/final_prof_igemm \
    python bench.py --steps 2 --warmup 3 $B --ncu-step 1 > $O/ncu2.log 2>&1; echo "ncu igemm rc=$?"
python tools/ncu_summary.py /tmp/final_prof_igemm.ncu-rep $O/final_ncu_igemm_summary.csv; mv $O/roofline_traffic.json $O/final_roofline_traffic.json 2>/dev/null
ls -la /tmp/final_prof_*.ncu-rep
;;
d)
python bench.py --steps 2 --warmup 3 $B > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -k 'regex:loss|router|hinge|expm1|disc_|gn_|adam|gen_out|ln_|bn' -c 40 -f -o /tmp/final_prof_hbm \
    python bench.py --steps 2 --warmup 3 $B --ncu-step 2 > $O/ncu3.log 2>&1; echo "ncu hbm rc=$?"
python tools/ncu_summary.py /tmp/final_prof_hbm.ncu-rep $O/final_ncu_hbm_summary.csv; rm -f $O/roofline_traffic.json
;;
e)
python tools/train_equivalence.py --out $O/final_train_equivalence.json > $O/final_train_equivalence.log 2>&1; echo "equivalence rc=$?"; tail -40 $O/final_train_equivalence.log
;;
esac
ls -la $O/final_* 2>/dev/null | awk '{print $5, $9}'
