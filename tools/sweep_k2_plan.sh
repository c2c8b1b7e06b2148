for tau in 2 3 4 5; do for mg in 24 64 128; do
echo "tau=$tau mg=$mg $(QC_K2_TAU=$tau QC_K2_MIN_GAIN=$mg python tools/microbench_k2.py 28 100000 tiles 2>&1 | head -1)"
done; done
