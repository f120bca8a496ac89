source tools/run_variants.sh
run base "" A=1
