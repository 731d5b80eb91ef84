# Round-2 profile set (one B200): plain run, ncu launch list, ncu --set full of the loop's kernels, block-width sweeps.
# The loop graph is switched off for the ncu passes so that every launch of an iteration is listed by itself.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-lobpcg --no-c5 --no-tight --c3-grid 0"
export DE_B200_LOOP_GRAPH=0
$B > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spmm_brb_kernel|ts2_update_kernel|ts2_gram_kernel|reduce_tail_kernel" -s 400 -c 10 -o gpurun_out/r02_full -f $B > gpurun_out/r02_ncu2.log 2>&1
unset DE_B200_LOOP_GRAPH
python tools/kernel_sweep.py --grid 100 --stencil q1 --csv gpurun_out/r02_sweep_q1100.csv > gpurun_out/r02_sweep_q1100.log 2>&1
python tools/kernel_sweep.py --grid 200 --stencil q1 --csv gpurun_out/r02_sweep_q1200.csv > gpurun_out/r02_sweep_q1200.log 2>&1
python tools/kernel_sweep.py --grid 200 --stencil fd --csv gpurun_out/r02_sweep_fd200.csv > gpurun_out/r02_sweep_fd200.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
ls -la gpurun_out | tail -14
