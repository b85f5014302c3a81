# Regenerates the round's evidence on a GPU box (run through gpurun; results land in gpurun_out/, the summaries worth keeping
# are copied to profiles/ by hand):   bash tools/refresh_profiles.sh r02
set -x
R=${1:-r02}
O=gpurun_out
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/${R}_bench_ref.json 2>> $O/${R}_bench_n1.err
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config --profiler-range > $O/${R}_launches_bench.log 2>&1
python tools/ncu_summary.py launches $O/${R}_launches.csv > $O/${R}_launches_bench.md
python tools/ncu_summary.py traffic $O/${R}_launches.csv > $O/${R}_bench_traffic.json
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fq_affine_sites_kernel -c 1 -f -o /tmp/${R}_dom python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config --profiler-range > $O/${R}_dom.log 2>&1
python tools/ncu_summary.py report /tmp/${R}_dom.ncu-rep > $O/${R}_fq_affine_sites.md
cp /tmp/${R}_dom.ncu-rep $O/
ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${R}_variants python tools/ncu_targets.py --order $O/${R}_variants_order.json > $O/${R}_ncu.log 2>&1
python tools/ncu_summary.py variants /tmp/${R}_variants.ncu-rep $O/${R}_variants_order.json > $O/${R}_variants_ncu_full.md
python tools/kbench.py --json $O/${R}_kbench.json > $O/${R}_kbench.log 2>&1
python tools/reference_on_gpu.py --json $O/${R}_reference_on_gpu.json > $O/${R}_reference_on_gpu.log 2>&1
python tools/config_bench.py --json $O/${R}_config_bench.json > $O/${R}_config_bench.log 2>&1
python tools/scale_bench.py --json $O/${R}_scale_n1.json > $O/${R}_scale_n1.log 2>&1
python tools/pcie_probe.py --json $O/${R}_pcie_probe.json > $O/${R}_pcie_probe.txt 2>&1
python tools/call_overhead.py > $O/${R}_call_overhead.txt 2>&1
tail -c 1500 $O/${R}_bench_n1.json
