for a in 0 1 3 4 7; do echo "ablate $a"; CB200_BWD_ABLATE=$a python tools/microbench.py 2>&1 | grep "attention bwd"; done
