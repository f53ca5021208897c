tag=$1
mkdir -p gpurun_out
cap() {
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/${tag}_$1" \
    python tools/prof_kernels.py "$4" ${5:-0} > "gpurun_out/${tag}_$1.log" 2>&1
  echo "$1 rc=$?"
}
for k in "$@"; do
  case $k in
    fwd) cap attn_fwd window_attn_fwd 2 attn_fwd ;;
    bwd) cap attn_bwd window_attn_bwd 2 attn_bwd ;;
    fwd2) cap attn_fwd_s2 window_attn_fwd 2 attn_fwd 1 ;;
    bwd2) cap attn_bwd_s2 window_attn_bwd 2 attn_bwd 1 ;;
  esac
done
