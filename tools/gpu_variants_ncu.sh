# usage: bash tools/gpu_variants_ncu.sh name1 name2 ...  -> instruction count / issue utilisation of k_photometric per variant
cd /root/repo
for n in "$@"; do
  L=/root/repo/variants/libvsl_$n.so; [ "$n" = "product" ] && L=/root/repo/unsupervised_pose_estimation_b200/libvsl_b200.so
  VSL_LIB_PATH=$L timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:k_photometric -s 2 -c 1 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-strong ${BENCH_ARGS} 2>/dev/null \
    | grep -E "k_photometric" | awk -F'","' -v n=$n '{print n, $(NF-2), $(NF)}' | tr -d '"'
done | tee -a gpurun_out/variants_ncu.txt
