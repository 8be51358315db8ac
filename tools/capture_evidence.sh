# Evidence run for profiles/: launch lists and --set full captures of the encoder GEMMs, the 256-row decode chain and the beam
# step, each ncu command only after the same command has exited 0 without ncu.  Run on the GPU box: bash tools/capture_evidence.sh
# (writes gpurun_out/r2c_*; summarise here with tools/launch_summary.py and tools/ncu_summary.py).
set -x
K='gemm_tcgen05|layernorm|vit_|preprocess|pool_prefix|rowstats|ln_stats|cls_rows'
python tools/prof_encoder.py 2 > gpurun_out/r2c_plain_enc.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$K" -s 88 -c 88 --csv --log-file gpurun_out/r2c_encoder_launches.csv python tools/prof_encoder.py 2 > gpurun_out/r2c_ncu1.log 2>&1
# bench.py launch list (2 timed steps after 3 warm-ups; first 2500 launches of our kernels)
python bench.py --steps 2 --warmup 3 --no-extra-configs --no-cpu-baseline > gpurun_out/r2c_plain_bench.json 2> gpurun_out/r2c_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2c_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extra-configs --no-cpu-baseline > gpurun_out/r2c_ncu2.log 2>&1
# full captures: GEMM (one per epilogue type) of the encoder
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 52 -c 5 -o gpurun_out/r2c_gemm python tools/prof_encoder.py 2 > gpurun_out/r2c_ncu3.log 2>&1
# the 256-row decode chain (tcgen05 64-wide tiles): one layer's kernels
python tools/prof_decode.py 256 > gpurun_out/r2c_plain_dec256.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|gpt_attention" -s 300 -c 6 -o gpurun_out/r2c_dec256 python tools/prof_decode.py 256 > gpurun_out/r2c_ncu4.log 2>&1
# beam step kernels
python tools/prof_beam.py > gpurun_out/r2c_plain_beam.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"beam_scores|beam_update|beam_merge" -s 12 -c 3 -o gpurun_out/r2c_beam python tools/prof_beam.py > gpurun_out/r2c_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
