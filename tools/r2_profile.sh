#!/bin/bash
# Round-2 evidence for profiles/: bench lines (default, bf16, 10 s x 24, 1 s x 128, widened, reference arm), latency sweep,
# ncu launch list and `--set full` captures.  Run on a B200: bash tools/r2_profile.sh ; outputs land in gpurun_out/.
mkdir -p gpurun_out
T=${1:-r2}
python bench.py --profile-out gpurun_out/${T}_bench_steps.json --latency-sweep gpurun_out/${T}_latency_sweep.json > gpurun_out/${T}_bench_line.json 2> gpurun_out/${T}_bench.err
python bench.py --precision bf16 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_line_bf16.json 2>> gpurun_out/${T}_bench.err
python bench.py --seconds 10 --batch 24 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_line_10s_b24.json 2>> gpurun_out/${T}_bench.err
python bench.py --seconds 1 --batch 128 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_line_1s_b128.json 2>> gpurun_out/${T}_bench.err
python bench.py --widened --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_line_widened.json 2>> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_reference_line.json 2>> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu1.log 2>&1
# `--set full` of one whole step (the second one; the first is the warm-up): the 34 GEMM / conv launches, then every other
# kernel.  The reports are condensed to CSV on the box (raw page -> tools/ncu_summary.py) and deleted: gpurun only
# brings back 64 MiB.
ncu --set full --clock-control none -k regex:"igemm" -s 34 -c 34 -f -o gpurun_out/${T}_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu2.log 2>&1
ncu -i gpurun_out/${T}_gemm.ncu-rep --page raw --csv > gpurun_out/${T}_gemm_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${T}_gemm_raw.csv gpurun_out/${T}_ncu_gemm_conv_full_summary.csv gpurun_out/${T}_traffic.json
rm -f gpurun_out/${T}_gemm.ncu-rep gpurun_out/${T}_gemm_raw.csv
ncu --set full --clock-control none -k regex:"^(?!.*igemm).*" -s 30 -c 30 -f -o gpurun_out/${T}_other python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${T}_ncu3.log 2>&1
ncu -i gpurun_out/${T}_other.ncu-rep --page raw --csv > gpurun_out/${T}_other_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${T}_other_raw.csv gpurun_out/${T}_ncu_other_kernels_full_summary.csv
rm -f gpurun_out/${T}_other.ncu-rep gpurun_out/${T}_other_raw.csv
tail -n 2 gpurun_out/${T}_ncu2.log gpurun_out/${T}_ncu3.log
ls -la gpurun_out | grep ${T}_
du -sh gpurun_out
