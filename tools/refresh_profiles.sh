set -x
O=gpurun_out
python bench.py > $O/r01d_bench_n1.json 2> $O/r01d_bench_n1.err
python bench.py --impl reference > $O/r01d_bench_ref.json 2>> $O/r01d_bench_n1.err
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r01d_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profiler-range > $O/r01d_launches_bench.log 2>&1
python tools/ncu_summary.py launches $O/r01d_launches.csv > $O/r01d_launches_bench.md
python tools/ncu_summary.py traffic $O/r01d_launches.csv > $O/r01d_bench_traffic.json
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fq_affine_kernel -s 3 -c 1 -f -o /tmp/r01d_dom python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profiler-range > $O/r01d_dom.log 2>&1
python tools/ncu_summary.py report /tmp/r01d_dom.ncu-rep > $O/r01d_fq_affine_f32_pt.md
ls -la /tmp/r01d_dom.ncu-rep; cp /tmp/r01d_dom.ncu-rep $O/ 
ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r01d_variants python tools/ncu_targets.py --order $O/r01d_variants_order.json > $O/r01d_ncu.log 2>&1
python tools/ncu_summary.py variants /tmp/r01d_variants.ncu-rep $O/r01d_variants_order.json > $O/r01d_variants_ncu_full.md
python tools/kbench.py --json $O/r01d_kbench.json > $O/r01d_kbench.log 2>&1
python tools/config_bench.py --json $O/r01d_config_bench.json > $O/r01d_config_bench.log 2>&1
python tools/scale_bench.py --json $O/r01d_scale_n1.json > $O/r01d_scale_n1.log 2>&1
python tools/pcie_probe.py > $O/r01d_pcie_probe.txt 2>&1
tail -c 1500 $O/r01d_bench_n1.json
ls -la $O
