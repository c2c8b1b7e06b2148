for tau in 3 5 7 9; do for mg in 48 128; do
echo "tau=$tau mg=$mg $(QC_K2_TAU=$tau QC_K2_MIN_GAIN=$mg python tools/microbench_k2_split.py 28 100000 all 2>&1 | head -1)"
done; done
