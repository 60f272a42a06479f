#!/bin/bash
# tools/gpu_session.sh <stage> — the ONE runner for GPU sessions (run under gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh tests'
# Every stage writes into gpurun_out/<tag>_*; copy what should be judged into profiles/ afterwards.
#   tests      pytest -m gpu (whole suite)                         [TAG, PYTEST_ARGS]
#   bench      python bench.py (N = 1)                             [TAG, BENCH_ARGS]
#   sweep      tools/sweep_order.py, select with and without sectors [TAG]
#   launches   ncu launch list of a short bench.py run             [TAG]
#   ncu        ncu --set full of the binned pipelines' kernels     [TAG]
#   all1       tests + bench + sweep + launches + ncu
#   multi      (gpurun --gpus N) multi-rank tests + torchrun bench at 1..N   [TAG, NGPU]
set -u
stage=${1:-all1}
TAG=${TAG:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run_tests() {
    timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > $OUT/${TAG}_pytest_gpu.txt 2>&1
    echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.txt; tail -5 $OUT/${TAG}_pytest_gpu.txt
}
run_bench() {
    timeout 900 python bench.py ${BENCH_ARGS:-} > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
    echo "bench rc=$?"; tail -c 3000 $OUT/${TAG}_bench_n1.json; tail -5 $OUT/${TAG}_bench_n1.err
}
run_sweep() {
    rm -f $OUT/${TAG}_sweep_order.jsonl
    timeout 600 python tools/sweep_order.py --tag pos --out $OUT/${TAG}_sweep_order.jsonl > $OUT/${TAG}_sweep.log 2>&1
    SDSLGPU_SELECT_SECTORS=0 timeout 600 python tools/sweep_order.py --tag sampled_select --ops select1 --out $OUT/${TAG}_sweep_order.jsonl >> $OUT/${TAG}_sweep.log 2>&1
    tail -30 $OUT/${TAG}_sweep.log
}
run_launches() {
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras ${BENCH_ARGS:-} > $OUT/${TAG}_launches_bench.log 2>&1
    python tools/summarize_launch_csv.py $OUT/${TAG}_launches_bench.csv > $OUT/${TAG}_launches_bench_summary.txt 2>&1; tail -30 $OUT/${TAG}_launches_bench_summary.txt
}
run_ncu() {
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'bin_' -s 12 -c 6 -f -o $OUT/${TAG}_ncu_binned \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/${TAG}_ncu_binned.log 2>&1
    python tools/summarize_ncu.py $OUT/${TAG}_ncu_binned.ncu-rep $OUT/${TAG}_ncu_full_binned.txt > /dev/null 2>&1; tail -50 $OUT/${TAG}_ncu_full_binned.txt
}
case $stage in
tests) run_tests ;;
bench) run_bench ;;
sweep) run_sweep ;;
launches) run_launches ;;
ncu) run_ncu ;;
all1) run_tests; run_bench; run_sweep; run_launches; run_ncu ;;
quick1) run_bench; run_sweep; run_launches; run_ncu ;;
multi)
    N=${NGPU:-2}
    timeout 900 python -m pytest tests/test_group_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest_multi.txt 2>&1; tail -15 $OUT/${TAG}_pytest_multi.txt
    nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
    : > $OUT/${TAG}_probe_pcie.jsonl
    for n in 1 2 4 8; do
        [ $n -gt $N ] && break
        if [ $n -eq 1 ]; then timeout 300 python tools/probe_pcie.py >> $OUT/${TAG}_probe_pcie.jsonl 2>/dev/null
        else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) tools/probe_pcie.py >> $OUT/${TAG}_probe_pcie.jsonl 2>/dev/null; fi
    done
    cat $OUT/${TAG}_probe_pcie.jsonl
    for n in 1 2 4 8; do
        [ $n -gt $N ] && break
        if [ $n -eq 1 ]; then
            timeout 900 python bench.py --gpus 1 ${BENCH_ARGS:-} ${BENCH_ARGS_N1:-} > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
        else
            timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
                bench.py --gpus $n ${BENCH_ARGS:-} > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
        fi
        echo "bench n=$n rc=$?"; tail -c 1500 $OUT/${TAG}_bench_n$n.json; tail -3 $OUT/${TAG}_bench_n$n.err
    done ;;
*) echo "unknown stage $stage"; exit 2 ;;
esac
