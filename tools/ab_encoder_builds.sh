# A/B of two builds of libvcb200.so on ONE box: per-kernel CUDA-event times of an encoder pass, builds alternated.
# usage: bash tools/ab_encoder_builds.sh path/to/other.so [rounds]
OTHER=$1; R=${2:-3}
for r in $(seq $R); do
  for which in other head; do
    if [ $which = other ]; then export VC_LIB=$OTHER; else unset VC_LIB; fi
    echo "== $which (round $r)"
    python tools/prof_encoder_kernels.py vit_b16_gpt2 64 16 | grep -E "gemm_lnf|gemm_resid|^sum"
  done
done
