# usage: bash tools/gpu_variants.sh "<probe args>" name1 name2 ...   (variants/libvsl_<name>.so; "product" = the in-tree library)
cd /root/repo
ARGS="$1"; shift
for n in "$@"; do
  if [ "$n" = "product" ]; then timeout 300 python tools/variant_probe.py $ARGS 2>/dev/null | tail -1
  else VSL_LIB_PATH=/root/repo/variants/libvsl_$n.so timeout 300 python tools/variant_probe.py $ARGS 2>/dev/null | tail -1; fi
done | tee -a gpurun_out/variants.jsonl
